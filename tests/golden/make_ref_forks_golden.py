"""Id-level pin of the Inmemory and Exactdistance storage modes against the reference's OWN forks.

The forks (BANG_Inmemory/parANN.cu, BANG_Exactdistance/parANN.cu) are stand-alone programs with N, D, MEDOID, the
element type, L and the chunk count compiled in; oracle/build_ref_forks.sh builds one binary per (fixture, L, chunks)
from a patched temporary copy (header block for the fixture, three empty BFS hooks, an empty Boost header, and one
added fwrite of the result ids the program already holds on the host).

    python tests/golden/make_ref_forks_golden.py build            # here: compiles every binary into oracle/_ref/
    gpurun -- 'python tests/golden/make_ref_forks_golden.py run'  # GPU box: writes gpurun_out/ref_forks_golden.npz
    cp gpurun_out/ref_forks_golden.npz tests/golden/              # commit

Keys: ids_<fork>_<case>_L<L>_rep<r> = u32[Q][k] (transposed from the program's [k][Q]), k = 10, 3 repetitions.
tests/test_oracle.py::test_oracle_reproduces_reference_forks compares the oracle with them on CPU.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import bang_b200  # noqa: E402,F401
from bang_b200 import formats  # noqa: E402

REFDIR = os.path.join(ROOT, "oracle", "_ref")
K = 10
# case -> (fixture npz, PQ key suffix, worklist lengths)
CASES = {
    "c1": ("fx_c1", "", (20, 64, 152)),          # C1: N = 10^4, D = 128 u8, m = 32
    "c1m128": ("fx_c1", "128", (64,)),           # the reference's SIFT1BSMALL chunk count (parANN.h:87)
    "f32": ("fx_f32", "", (32,)),
    "i8": ("fx_i8", "", (32,)),
}
CT = {"uint8": "uint8_t", "int8": "int8_t", "float": "float"}


def load_case(case):
    name, suf, Ls = CASES[case]
    z = np.load(os.path.join(HERE, name + ".npz"))
    return z, suf, Ls


def build():
    for case in CASES:
        z, suf, Ls = load_case(case)
        N, D = z["base"].shape
        m = z["codes" + suf].shape[1]
        for L in Ls:
            subprocess.run([os.path.join(ROOT, "oracle", "build_ref_forks.sh"), case, CT[formats.dtype_name(z["base"])], str(N),
                            str(D), str(int(z["medoid"])), str(L), str(m)], check=True)


def run():
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for case in CASES:
            z, suf, Ls = load_case(case)
            prefix = os.path.join(tmp, case)
            p = formats.write_index(prefix, z["base"], z["deg"], z["nbrs"], int(z["medoid"]), z["pivots" + suf],
                                    z["centroid" + suf], z["chunk_offsets" + suf], z["codes" + suf])
            formats.write_bin(p.query, z["queries"])
            formats.write_truthset(p.truth, z["gt_ids"], z["gt_dists"])
            Q = len(z["queries"])
            for L in Ls:
                for fork in ("inmem", "exact"):
                    exe = os.path.join(REFDIR, f"bang_{fork}_{case}_L{L}")
                    for rep in range(3):
                        ids_path = os.path.join(tmp, "ids.bin")
                        if os.path.exists(ids_path):
                            os.remove(ids_path)
                        r = subprocess.run([exe, p.old_pivots, p.pq_compressed, p.disk, p.query, p.old_chunk_offsets, p.old_centroid,
                                            p.truth, str(Q), "1", "256", "512", "256", str(K), "8", "0"], input="n\n",
                                           capture_output=True, text=True, timeout=600, env=dict(os.environ, BANG_DUMP_IDS=ids_path))
                        if r.returncode != 0 or not os.path.exists(ids_path):
                            print(r.stdout[-3000:], r.stderr[-3000:])
                            raise SystemExit(f"{exe} failed")
                        out[f"ids_{fork}_{case}_L{L}_rep{rep}"] = np.fromfile(ids_path, dtype=np.uint32).reshape(K, Q).T.copy()
                        if rep == 0:
                            tail = [ln for ln in r.stdout.splitlines() if ln.strip()][-4:]
                            print(fork, case, L, "|", " / ".join(tail), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "ref_forks_golden.npz"), **out)
    print("wrote gpurun_out/ref_forks_golden.npz with", len(out), "arrays")


if __name__ == "__main__":
    {"build": build, "run": run}[sys.argv[1]]()
