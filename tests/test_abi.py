"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every function include/kmernator_b200.h
declares, fills the reference's option defaults, and fails loudly (no CPU fallback) when there is no CUDA device.
Also the JSON contract of the reference arm of bench.py (the one CPU leg that may execute oracle/)."""
import ctypes as C
import json
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "kmernator_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kmn_[a-z0-9_]+)\s*\(", src)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    import kmernator_b200 as K
    lib = K.capi.load()
    names = _declared_functions()
    assert len(names) >= 20 and "kmn_count_batch" in names and "kmn_trim_batch" in names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.kmn_version.restype = C.c_char_p
    assert b"sm_100a" in lib.kmn_version()


def test_default_opts_are_the_reference_defaults():
    import kmernator_b200 as K
    lib = K.capi.load()
    o = K.capi.KmnOpts()
    lib.kmn_default_opts(C.byref(o))
    assert o.struct_size == C.sizeof(K.capi.KmnOpts)
    assert o.min_quality_score == 3                   # src/Options.h:329
    assert abs(o.min_kmer_quality - 0.10) < 1e-7      # src/KmerSpectrum.h:92
    assert o.min_depth == 2                           # src/KmerSpectrum.h:92
    assert o.fastq_start_char == 33 and o.hash_kind == 0 and o.value_kind == 0


def test_invalid_options_are_rejected_before_touching_a_device():
    import kmernator_b200 as K
    for kw in (dict(kmer_size=0), dict(kmer_size=129), dict(fastq_start_char=50)):
        with pytest.raises(K.capi.KmnError):
            K.Context(**kw)


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_a_gpu():
    import kmernator_b200 as K
    with pytest.raises(K.capi.KmnError) as e:
        K.Context(kmer_size=31)
    assert "no CPU fallback" in str(e.value)


def test_reference_arm_contract():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-reads", "20000"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                             # ONE JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "kmers/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "kmers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"]
