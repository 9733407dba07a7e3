"""pattern/sharded.h::shard_bounds (the --num-gpus split of the drop-in binaries) on the CPU: contiguous, complete,
16-element aligned shard starts, and identical to the helper the 2-GPU tests use (tests/test_comm_gpu.py::_bounds)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROGRAM = r"""
#include "sharded.h"
#include <cstdio>
int main() {
  const std::size_t sizes[] = {0, 1, 15, 16, 17, 1000, 4099, 1u << 20, (1u << 22) + 3, 268435456ull, 8589934592ull};
  for(std::size_t n : sizes)
    for(int world : {1, 2, 3, 4, 8})
      for(int r = 0; r < world; ++r) {
        const auto b = pattern::shard_bounds(n, world, r);
        std::printf("%zu %d %d %zu %zu\n", n, world, r, b.first, b.second);
      }
  return 0;
}
"""


def _bounds(n, world, rank):  # same as tests/test_comm_gpu.py
    per = -(-n // world)
    per = -(-per // 16) * 16
    b = min(n, rank * per)
    return b, min(n, b + per)


def test_cpp_shard_bounds(tmp_path):
    src = tmp_path / "bounds.cpp"
    exe = tmp_path / "bounds"
    src.write_text(PROGRAM)
    pkg = os.path.join(ROOT, "sycl-bench_b200")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", f"-I{pkg}/pattern", f"-I{pkg}/harness", f"-I{ROOT}/include", str(src), "-o", str(exe),
                    f"-L{pkg}/lib", "-lsb_cuda", f"-Wl,-rpath,{pkg}/lib", "-ldl", "-lpthread"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    rows = [tuple(int(v) for v in line.split()) for line in out.splitlines()]
    assert len(rows) == 11 * (1 + 2 + 3 + 4 + 8)
    prev_end = {}
    for n, world, r, b, e in rows:
        assert (b, e) == _bounds(n, world, r), (n, world, r)
        assert b <= e <= n and (b % 16 == 0 or b == n)
        if r == 0:
            assert b == 0
        else:
            assert b == prev_end[(n, world)]  # contiguous
        prev_end[(n, world)] = e
        if r == world - 1:
            assert e == n  # complete
