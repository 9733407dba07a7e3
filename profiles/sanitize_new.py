"""Small exerciser for compute-sanitizer (memcheck / racecheck / synccheck) over the kernels added late in round 1:
sb_fill, sb_copy, sb_dot, sb_linreg, the 18- / 15-warp scan instantiations and (SB_SCAN_TB=1) scan_tb_kernel.

    compute-sanitizer --tool memcheck python profiles/sanitize_new.py
    SB_SCAN_TB=1 compute-sanitizer --tool racecheck python profiles/sanitize_new.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
ctx = pkg.Context(0, torch.cuda.current_stream().cuda_stream)
TD = {"i32": torch.int32, "i64": torch.int64, "f32": torch.float32, "f64": torch.float64}
n = (1 << 20) + 77
for t, dt in TD.items():
    x = torch.empty(n + 8, dtype=dt, device="cuda")
    y = torch.empty(n + 8, dtype=dt, device="cuda")
    out = torch.zeros(8, dtype=dt, device="cuda")
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    ctx.fill(x, 1, n=n)
    ctx.fill(y[1:], 2, n=n)                       # element-aligned only
    ctx.copy(y[4:], x[4:], (n - 8) * x.element_size())
    ctx.copy(y[1:], x[2:], 1001)                  # non-congruent pointers: byte path
    ctx.dot(x, y, out, n=n)                       # vector path
    ctx.dot(x[1:], y[1:], out, n=n - 1)           # scalar path
    if t[0] == "f":
        ctx.linreg(x, y, out, n=n)
        ctx.linreg(x[1:], y[1:], out, n=n - 1)
    ctx.iota(x, n=n)
    ctx.scan(x, y, n=n)                           # global scan: 18-warp look-back kernel (or scan_tb_kernel)
    ctx.scan(x, y, exclusive=True, n=n)
    if t[0] == "i":
        ctx.check_iota(y, bad, kind=1, n=n)
    ctx.scan(x, y, exclusive=True, group_len=256, n=n)   # independent tiles: 18 / 15 warps
    ctx.segscan(x, y, seg_len=1 << 16, n=n - 77)         # segmented look-back, 16 warps
    ctx.sync()
    if t[0] == "i":
        assert int(bad.cpu()[0]) == 0, t
comm = pkg.Comm(ctx, 1)
comm.close()
ctx.close()
print("sanitize_new: done, tb =", os.environ.get("SB_SCAN_TB", "0"))
