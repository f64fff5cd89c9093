"""Build librc_b200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles).

    python -m raincast_gnn_b200.csrc.build [--force]

The library is a plain C-ABI shared object (include/rc_b200.h); it links only the CUDA runtime.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SOURCES = ["rc_core.cu", "rc_graph.cu", "rc_gine.cu", "rc_gine_wide.cu", "rc_gine_tiled.cu", "rc_tiles.cu", "rc_gemm.cu", "rc_gemm_tc.cu", "rc_deepsets.cu", "rc_deepsets_tc.cu", "rc_deepsets_tc_bwd.cu", "rc_crps.cu", "rc_head.cu", "rc_misc.cu", "rc_p2p.cu", "rc_debug.cu"]
HEADERS = ["rc_common.cuh", "rc_umma.cuh", "rc_crps_node.cuh", "rc_gemm_tile.cuh", "rc_gine_tile.cuh", "rc_deepsets_tile.cuh",
           "rc_crps_tile.cuh", "rc_misc_tile.cuh", os.path.join(ROOT, "include", "rc_b200.h")]
LIB = os.path.join(HERE, "librc_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
              "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-I", os.path.join(ROOT, "include"), "-I", HERE]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "nvcc")
    hdrs = [h if os.path.isabs(h) else os.path.join(HERE, h) for h in HEADERS]
    objs, jobs = [], []
    for src in SOURCES:
        s = os.path.join(HERE, src)
        o = os.path.join(HERE, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as pool:
            for res in pool.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs):
                if verbose or res.returncode != 0:
                    sys.stderr.write(res.stdout + res.stderr)
                if res.returncode != 0:
                    raise RuntimeError("nvcc failed: " + " ".join(res.args))
    if jobs or force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
