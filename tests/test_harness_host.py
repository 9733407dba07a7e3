"""Host-only pieces of the C++ harness, exercised on the CPU through a small driver program: the statistics and
column set of TimeMetricsProcessor (reference: include/time_metrics.h:56-147), the CSV writer
(include/result_consumer.h:54-106) and the command-line parser (include/command_line.h:57-127)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROGRAM = r"""
#include "common.h"
#include <cstdio>
struct WithMetric {
  static ThroughputMetric getThroughputMetric(const BenchmarkArgs&) { return {2.0, "GiB"}; }
};
struct WithoutMetric {};
struct Printer : ResultConsumer {
  void proceedToBenchmark(const std::string& n) override { std::printf("BENCH %s\n", n.c_str()); }
  void consumeResult(const std::string& k, const std::string& v, const std::string& u = "") override { std::printf("%s=%s [%s]\n", k.c_str(), v.c_str(), u.c_str()); }
  void flush() override {}
};
int main(int argc, char** argv) {
  using ns = std::chrono::nanoseconds;
  Printer out;
  {
    BenchmarkArgs a;  // no warm-up run
    TimeMetricsProcessor<WithMetric> m(a);
    for(long t : {5000000L, 1000000L, 3000000L, 2000000L}) m.addTimingResult("kernel-time", ns(t));
    m.markAsUnavailable("submit-time");
    out.proceedToBenchmark("plain");
    m.emitResults(out);
  }
  {
    BenchmarkArgs a;
    a.warmup_run = true;  // the first sample is dropped
    TimeMetricsProcessor<WithoutMetric> m(a);
    for(long t : {9000000L, 1000000L, 3000000L}) m.addTimingResult("run-time", ns(t));
    out.proceedToBenchmark("warm");
    m.emitResults(out);
  }
  {
    AppendingCsvResultConsumer csv(argv[1]);
    csv.proceedToBenchmark("B1");
    csv.consumeResult("zeta", "1");
    csv.consumeResult("alpha", "2", "s");
    csv.flush();
    csv.proceedToBenchmark("B2");
    csv.consumeResult("alpha", "3");
    csv.discard();  // a failed benchmark leaves no row
    csv.proceedToBenchmark("B3");
    csv.consumeResult("beta", "4");
    csv.flush();
  }
  {
    const char* args[] = {"prog", "--size=4096", "--flag", "--verification-begin=1,2,3", "--name=x=y"};
    CommandLine cli(5, const_cast<char**>(args));
    const auto id = cli.get<sb::id3>("--verification-begin");
    std::printf("CLI %zu %d %d %zu,%zu,%zu %s\n", cli.getOrDefault<std::size_t>("--size", 1), (int)cli.isFlagSet("--flag"), (int)cli.isFlagSet("--other"), id[0], id[1],
        id[2], cli.get<std::string>("--name").c_str());
    try {
      const char* dup[] = {"prog", "--size=1", "--size=2"};
      CommandLine bad(3, const_cast<char**>(dup));
      std::printf("DUP accepted\n");
    } catch(const std::invalid_argument& e) { std::printf("DUP %s\n", e.what()); }
    try {
      cli.get<int>("--missing");
    } catch(const std::invalid_argument& e) { std::printf("MISSING %s\n", e.what()); }
  }
  return 0;
}
"""


@pytest.fixture(scope="module")
def driver_output(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("harness_host")
    src, exe, csv = tmp / "driver.cpp", tmp / "driver", tmp / "out.csv"
    src.write_text(PROGRAM)
    pkg = os.path.join(ROOT, "sycl-bench_b200")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", f"-I{pkg}/harness", f"-I{ROOT}/include", str(src), "-o", str(exe),
                    f"-L{pkg}/lib", "-lsb_cuda", f"-Wl,-rpath,{pkg}/lib", "-ldl", "-lpthread"], check=True)
    out = subprocess.run([str(exe), str(csv)], capture_output=True, text=True, check=True).stdout
    return out, csv.read_text()


def _block(out, name):
    lines = out.split(f"BENCH {name}\n", 1)[1].split("BENCH ", 1)[0].split("CLI ", 1)[0].splitlines()
    return dict(line.split("=", 1) for line in lines if "=" in line)


def test_time_metrics_statistics(driver_output):
    d = _block(driver_output[0], "plain")
    # samples 1, 2, 3, 5 ms: mean 2.75 ms, SAMPLE standard deviation, UPPER median, min, throughput = metric / min
    assert d["kernel-time-mean"] == "0.002750 [s]"
    assert d["kernel-time-stddev"] == "0.001708 [s]"          # sqrt(8.75e-6 / 3)
    assert d["kernel-time-median"] == "0.003000 [s]"          # element n/2 of the sorted samples
    assert d["kernel-time-min"] == "0.001000 [s]"
    assert d["kernel-time-samples"] == '"0.001000 0.002000 0.003000 0.005000" []'
    assert d["kernel-time-throughput"] == "2000.000000 [GiB/s]"
    assert d["kernel-time-min-ns"] == "1000000 [ns]"
    assert d["throughput-metric"] == "2.000000 [GiB]"
    for s in ("mean", "stddev", "median", "min", "samples", "throughput"):
        assert d[f"submit-time-{s}"] == "N/A []"               # unavailable timings still emit their columns


def test_warmup_run_drops_the_first_sample_and_missing_metric_is_na(driver_output):
    d = _block(driver_output[0], "warm")
    assert d["run-time-mean"] == "0.002000 [s]" and d["run-time-min"] == "0.001000 [s]"
    assert d["run-time-samples"] == '"0.001000 0.003000" []'
    assert d["throughput-metric"] == "N/A []" and d["run-time-throughput"] == "N/A []"


def test_csv_writer_format(driver_output):
    lines = driver_output[1].splitlines()
    # header re-emitted per benchmark, columns sorted, discarded benchmark leaves nothing
    assert lines == ["# Benchmark name,alpha,zeta", "B1,2,1", "# Benchmark name,beta", "B3,4"]


def test_command_line_parser(driver_output):
    out = driver_output[0]
    assert "CLI 4096 1 0 1,2,3 x=y" in out
    assert "DUP Encountered command line argument several times: --size" in out
    assert "MISSING Command line argument was requested but missing: --missing" in out
