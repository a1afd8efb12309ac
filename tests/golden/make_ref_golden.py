"""Runs the UNMODIFIED reference CUDA build (oracle/_ref, see oracle/build_ref.sh) on a GPU box over the
committed fixtures and records what it returns.  The output pins the oracle against the real
reference (tests/test_oracle_golden.py, CPU-only):

    gpurun -- 'python tests/golden/make_ref_golden.py'        # writes gpurun_out/ref_golden.npz
    cp gpurun_out/ref_golden.npz tests/golden/ref_golden.npz  # commit

Keys: ids_<fixture>_L<L> = u64[Q][k] returned by BANGSearch<T>::bang_query (k = 10), 3 repetitions
(`rep<r>`) to expose run-to-run nondeterminism of the reference.
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

REF = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
K = 10
LS = (10, 32, 100)


def main():
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name in ("fx_u8", "fx_f32", "fx_i8"):
            z = np.load(os.path.join(HERE, name + ".npz"))
            prefix = os.path.join(tmp, name)
            p = formats.write_index(prefix, z["base"], z["deg"], z["nbrs"], int(z["medoid"]), z["pivots"], z["centroid"],
                                    z["chunk_offsets"], z["codes"])
            formats.write_bin(p.query, z["queries"])
            Q = len(z["queries"])
            dt = formats.dtype_name(z["base"])
            for L in LS:
                for rep in range(3):
                    ids_path = os.path.join(tmp, "ids.bin")
                    r = subprocess.run([REF, prefix, p.query, str(Q), str(K), str(L), dt, ids_path], capture_output=True,
                                       text=True, timeout=600)
                    if r.returncode != 0:
                        print(r.stdout[-3000:], r.stderr[-3000:])
                        raise SystemExit(f"ref_driver failed for {name} L={L}")
                    out[f"ids_{name}_L{L}_rep{rep}"] = np.fromfile(ids_path, dtype=np.uint64).reshape(Q, K)
                print(name, L, "ok")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "ref_golden.npz"), **out)
    print("wrote gpurun_out/ref_golden.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
