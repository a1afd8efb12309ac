"""Host-side pieces of bench.py that the driver's records depend on (no GPU): the reference arm's JSON line, the traffic lookup
(an ncu capture counts only for the exact workload / mode / L / batch / kernel source it was taken on), the option parser."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_traffic_lookup_is_keyed_by_the_kernel_source():
    b = _bench()
    entries = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    assert entries, "profiles/traffic.json holds the ncu DRAM bytes of the default workload"
    for e in entries:
        for key in ("source", "workload", "mode", "L", "queries", "dram_bytes_read", "dram_bytes_write", "kernel_source_hash"):
            assert key in e
        got = b.ncu_traffic(e["workload"], e["mode"], e["L"], e["queries"])
        if e["kernel_source_hash"] == b.kernel_source_hash():
            # (several captures of one key may exist; the first one wins)
            first = next(x for x in entries if (x["workload"], x["mode"], x["L"], x["queries"], x["kernel_source_hash"]) ==
                         (e["workload"], e["mode"], e["L"], e["queries"], e["kernel_source_hash"]))
            assert got == first["dram_bytes_read"] + first["dram_bytes_write"]
    assert b.ncu_traffic("deep100m", "inmemory", 7, 10000) is None        # an L nobody captured
    assert b.ncu_traffic("no-such-workload", "inmemory", 36, 10000) is None


def test_default_workload_is_the_largest_single_gpu_config():
    b = _bench()
    wl = b.WORKLOADS[b.DEFAULT_WORKLOAD]
    assert b.DEFAULT_WORKLOAD == "deep100m" and wl["n"] == 100_000_000 and wl["d"] == 96 and wl["q"] == 10_000
    assert b.index_bytes(wl) < 180 * 2**30                                   # fits one B200
    assert b.WORKLOADS["sift1b"]["sharded"] and b.WORKLOADS["sift256m"]["sharded"]


def test_points_alias_survives_torchrun_style_parsing():
    """torch.distributed.run's own parser rejects `--n` as ambiguous; bench.py also answers to --points."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True)
    assert out.returncode == 0 and "--points" in out.stdout and "--impl" in out.stdout
