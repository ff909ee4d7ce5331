"""Host-side helpers of bench.py that need no GPU: the parser of the same-box reference-GPU figure (reference arm) and the
workload description both arms share."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("zp_bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_parse_reference_gpu_log():
    b = _bench()
    text = ("debug noise\nreference gen_proof call 0: 9.303 s\nreference gen_proof call 1: 10.356 s\n"
            "reference gen_proof call 2: 6.993 s\nreference proof written; equals the oracle's proof: True\n")
    calls, same = b.parse_reference_gpu_log(text)
    assert calls == [9.303, 10.356, 6.993] and same is True
    calls, same = b.parse_reference_gpu_log("Segmentation fault\n")
    assert calls == [] and same is None


def test_reference_gpu_unavailable_without_device_or_library(monkeypatch):
    b = _bench()
    r = b.reference_native_gpu(4, repeat=1, timeout_s=30)
    # this container has no GPU (and the GPU box runs the real thing in the reference arm): never raises, says why
    assert "unavailable" in r or "calls_s" in r


def test_both_arms_describe_the_same_workload():
    b = _bench()
    a = b.workload_config(15, 3161924, 22, 1)
    assert "HEIGHT=15" in a["workload"] and "cs.n=3161924" in a["workload"] and "2^22" in a["workload"]
