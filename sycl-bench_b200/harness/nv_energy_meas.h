// NVML energy / clock hook — the reference reserves include/nv_energy_meas.h for this and leaves it empty
// (a one-line TODO); BenchmarkHook (include/benchmark_hook.h:6-16) is the interface it was meant to implement.
//
// Per benchmark row it adds:
//   energy-mj-mean     mean energy per run from NVML's total-energy counter (millijoule; the counter ticks every
//                      few milliseconds, so short kernels need many --num-runs for a meaningful value)
//   power-w            board power right after the last run
//   sm-clock-mhz       SM clock right after the last run, and sm-clock-max-mhz
//   throttle-reasons   names of the clock-event reasons seen after any run ("none" if none)
// NVML is loaded with dlopen("libnvidia-ml.so.1"); when it is not available the hook adds nothing.  The device is
// matched by PCI bus id (sb_ctx_pci_bus_id), because NVML indices ignore CUDA_VISIBLE_DEVICES.
//
// Included by sb_harness.h (common.h) right after BenchmarkHook is defined; include common.h, not this file.
#pragma once

#include <dlfcn.h>

#include <string>

class NvmlEnergyHook : public BenchmarkHook {
public:
  explicit NvmlEnergyHook(sb_ctx* ctx) {
    lib_ = dlopen("libnvidia-ml.so.1", RTLD_NOW | RTLD_LOCAL);
    if(!lib_) return;
    auto init = sym<int (*)()>("nvmlInit_v2");
    auto by_pci = sym<int (*)(const char*, void**)>("nvmlDeviceGetHandleByPciBusId_v2");
    energy_ = sym<int (*)(void*, unsigned long long*)>("nvmlDeviceGetTotalEnergyConsumption");
    power_ = sym<int (*)(void*, unsigned int*)>("nvmlDeviceGetPowerUsage");
    clock_ = sym<int (*)(void*, int, unsigned int*)>("nvmlDeviceGetClockInfo");
    max_clock_ = sym<int (*)(void*, int, unsigned int*)>("nvmlDeviceGetMaxClockInfo");
    reasons_ = sym<int (*)(void*, unsigned long long*)>("nvmlDeviceGetCurrentClocksEventReasons");
    if(!reasons_) reasons_ = sym<int (*)(void*, unsigned long long*)>("nvmlDeviceGetCurrentClocksThrottleReasons");
    char bus[64] = {0};
    if(!init || !by_pci || sb_ctx_pci_bus_id(ctx, bus, sizeof(bus)) != SB_OK || init() != 0 || by_pci(bus, &dev_) != 0) dev_ = nullptr;
  }
  ~NvmlEnergyHook() override {
    if(lib_) {
      if(auto shutdown = sym<int (*)()>("nvmlShutdown")) shutdown();
      dlclose(lib_);
    }
  }
  void atInit() override {}
  void preSetup() override {}
  void postSetup() override {}
  void preKernel() override {
    if(dev_ && energy_ && energy_(dev_, &e0_) != 0) e0_ = 0;
  }
  void postKernel() override {
    if(!dev_) return;
    unsigned long long e1 = 0;
    if(energy_ && e0_ && energy_(dev_, &e1) == 0 && e1 >= e0_) {
      energy_sum_mj_ += static_cast<double>(e1 - e0_);
      ++energy_runs_;
    }
    unsigned long long r = 0;
    if(reasons_ && reasons_(dev_, &r) == 0) seen_reasons_ |= r;
    if(power_) power_(dev_, &power_mw_);
    if(clock_) clock_(dev_, 1 /* NVML_CLOCK_SM */, &sm_mhz_);
  }
  void emitResults(ResultConsumer& rc) override {
    if(!dev_) return;
    if(energy_runs_) rc.consumeResult("energy-mj-mean", std::to_string(energy_sum_mj_ / energy_runs_), "mJ");
    if(power_) rc.consumeResult("power-w", std::to_string(power_mw_ / 1000.0), "W");
    if(clock_) rc.consumeResult("sm-clock-mhz", std::to_string(sm_mhz_), "MHz");
    unsigned int mx = 0;
    if(max_clock_ && max_clock_(dev_, 1, &mx) == 0) rc.consumeResult("sm-clock-max-mhz", std::to_string(mx), "MHz");
    if(reasons_) rc.consumeResult("throttle-reasons", reason_names(seen_reasons_));
  }

  // nvml.h: nvmlClocksEventReason* bit masks
  static std::string reason_names(unsigned long long r) {
    static const struct { unsigned long long bit; const char* name; } table[] = {
        {0x1ull, "gpu_idle"}, {0x2ull, "applications_clocks_setting"}, {0x4ull, "sw_power_cap"}, {0x8ull, "hw_slowdown"},
        {0x10ull, "sync_boost"}, {0x20ull, "sw_thermal_slowdown"}, {0x40ull, "hw_thermal_slowdown"},
        {0x80ull, "hw_power_brake_slowdown"}, {0x100ull, "display_clock_setting"}};
    std::string out;
    for(const auto& e : table)
      if(r & e.bit) out += (out.empty() ? "" : "|") + std::string(e.name);
    return out.empty() ? "none" : out;
  }

private:
  template <class F>
  F sym(const char* name) const { return lib_ ? reinterpret_cast<F>(dlsym(lib_, name)) : nullptr; }

  void* lib_ = nullptr;
  void* dev_ = nullptr;
  int (*energy_)(void*, unsigned long long*) = nullptr;
  int (*power_)(void*, unsigned int*) = nullptr;
  int (*clock_)(void*, int, unsigned int*) = nullptr;
  int (*max_clock_)(void*, int, unsigned int*) = nullptr;
  int (*reasons_)(void*, unsigned long long*) = nullptr;
  unsigned long long e0_ = 0, seen_reasons_ = 0;
  double energy_sum_mj_ = 0.0;
  unsigned long long energy_runs_ = 0;
  unsigned int power_mw_ = 0, sm_mhz_ = 0;
};
