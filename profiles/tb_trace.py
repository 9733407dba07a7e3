"""Phase trace of scan_tb_kernel: per-tile %globaltimer stamps of the CTA roles.

    make -C sycl-bench_b200 trace        # lib/tune/libsb_cuda_trace.so, sb_scan.cu compiled with -DSB_TB_TRACE
    SB_CUDA_LIB=sycl-bench_b200/lib/tune/libsb_cuda_trace.so python profiles/tb_trace.py [dtype] [log2n]
"""
import ctypes as C
import os
import sys

os.environ.setdefault("SB_SCAN_TB", "1")  # the traced kernel is behind this switch (read when the library loads)

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

EV = ["load_issue", "agg_in", "lb_done", "stored", "stage_free", "A_wait", "A_data", "A_done", "B_scan_done", "B_prefix", "B_done"]


def main():
    dtype = sys.argv[1] if len(sys.argv) > 1 else "i32"
    log2n = int(sys.argv[2]) if len(sys.argv) > 2 else 28
    pkg = ge.load_package()
    tdt = {"i32": torch.int32, "i64": torch.int64, "f32": torch.float32, "f64": torch.float64}[dtype]
    ctx = pkg.Context(0, torch.cuda.current_stream().cuda_stream)
    n = 1 << log2n
    x = torch.empty(n, dtype=tdt, device="cuda")
    y = torch.empty_like(x)
    ctx.iota(x)
    for _ in range(3):
        ctx.scan(x, y)
    torch.cuda.synchronize()
    L = pkg.lib()
    buf = np.zeros((160, 48, 16), dtype=np.uint64)
    L.sb_tb_trace_dump.argtypes = [C.c_void_p]
    assert L.sb_tb_trace_dump(buf.ctypes.data) == 0
    t = buf.astype(np.int64)
    for cta in (0, 73, 147):
        base = t[cta, 0, 0]
        print(f"== CTA {cta}: times in us relative to the first load issue; one line per tile")
        print("tile " + " ".join(f"{e:>11s}" for e in EV))
        for k in range(4, 20):
            print(f"{k:4d} " + " ".join(f"{(t[cta, k, e] - base) / 1000.0:11.2f}" for e in range(len(EV))))
    # aggregate: mean durations over CTAs and tiles 8..40
    sl = t[:148, 8:40, :]
    def d(a, b):
        return float((sl[:, :, a] - sl[:, :, b]).mean()) / 1000.0
    print("mean us: load latency (A_data - load_issue) %.2f | A wait for data %.2f | A compute %.2f | look-back (lb_done - agg_in) %.2f | B wait for prefix %.2f | B tail %.2f | store read %.2f"
          % (d(6, 0), d(6, 5), d(7, 6), d(2, 1), d(9, 8), d(10, 9), d(4, 3)))
    # per round: when was the last aggregate of the round in, who was last, when did the round prefix go out
    t0 = t[:148, 0, 0].min()
    print("round  first_agg   last_agg (cta)  rdesc_out(147)  first_prefix  last_prefix (cta)  last_B_done (cta)")
    for r in range(6, 22):
        agg = t[:148, r, 1] - t0
        lb = t[:148, r, 2] - t0
        bd = t[:148, r, 10] - t0
        print(f"{r:5d} {agg.min() / 1e3:10.2f} {agg.max() / 1e3:10.2f} ({int(agg.argmax()):3d}) {lb[147] / 1e3:14.2f} {lb.min() / 1e3:13.2f} {lb.max() / 1e3:12.2f} ({int(lb.argmax()):3d}) {bd.max() / 1e3:12.2f} ({int(bd.argmax()):3d})")
    per = (t[:148, 40, 10] - t[:148, 8, 10]).mean() / 32 / 1000.0
    print("mean period per tile %.2f us" % per)


if __name__ == "__main__":
    main()
