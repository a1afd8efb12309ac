"""Golden vectors for the DiskANN .index -> BANG .bin converter: runs the REFERENCE's own script
(/root/reference/BANG_Base/bang_preprocess.py, unmodified, as a subprocess) on small synthetic .index files and
stores input + outputs in tests/golden/preprocess_golden.npz.  Run in the build container (the reference is not
present on the GPU box):  python tests/golden/make_preprocess_golden.py"""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bang_b200  # noqa: F401
from bang_b200 import formats

SCRIPT = "/root/reference/BANG_Base/bang_preprocess.py"
CASES = [("u8", "uint8", 300, 16, 8), ("f32", "float", 57, 12, 6), ("i8", "int8", 120, 20, 5)]
out = {}
rng = np.random.default_rng(0xD15C)
with tempfile.TemporaryDirectory() as tmp:
    for name, dtype, n, d, R in CASES:
        if dtype == "float":
            vec = rng.normal(size=(n, d)).astype(np.float32)
        elif dtype == "uint8":
            vec = rng.integers(0, 256, size=(n, d), dtype=np.uint8)
        else:
            vec = rng.integers(-128, 128, size=(n, d), dtype=np.int8)
        deg = rng.integers(1, R + 1, size=n).astype(np.uint32)
        nbrs = rng.integers(0, n, size=(n, R)).astype(np.uint32)      # unsorted on purpose
        idx = os.path.join(tmp, f"{name}_disk.index")
        dst = os.path.join(tmp, f"{name}_disk.bin")
        formats.write_diskann_index(idx, vec, deg, nbrs, medoid=n // 3, rng=rng)
        r = subprocess.run([sys.executable, SCRIPT, idx, dst, str(d), str(formats.DTYPE_CODE[dtype]), str(R)], capture_output=True, text=True)
        assert r.returncode == 0 and "Total # of Nodes Discovered" in r.stdout, r.stdout + r.stderr
        out[f"{name}_index"] = np.fromfile(idx, dtype=np.uint8)
        out[f"{name}_bin"] = np.fromfile(dst, dtype=np.uint8)
        out[f"{name}_meta"] = np.fromfile(dst[:-4] + "_metadata.bin", dtype=np.uint8)
        out[f"{name}_args"] = np.array([d, formats.DTYPE_CODE[dtype], R, n], dtype=np.int64)
        print(name, "index", out[f"{name}_index"].size, "bin", out[f"{name}_bin"].size, "meta", out[f"{name}_meta"].size)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "preprocess_golden.npz"), **out)
