"""Pins oracle/pixel_oracle.py — the checker of the CUDA point<->pixel kernels (tests/test_gpu_pixel.py) — to
the reference's own code, run here on CPU (tests/pixel_pin_run.py in a fresh interpreter)."""
import json
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.timeout(900)
@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "core", "models", "fusion_blocks.py")), reason="reference tree not present (GPU box)")
def test_pixel_oracle_is_bit_identical_to_the_reference_code():
    """Feature_Gather / Feature_Fetch against core/models/fusion_blocks.py:241-278 called directly; the multi-scale
    point->pixel scatter-mean against the loop inside the UNMODIFIED student model's forward
    (spvcnn_swiftnet18_spformer_tsd_full.py:448-478), captured with forward pre-hooks at all four stages (grids 24x40 ... 3x5,
    32 ... 256 channels, a blind camera included).  Same torch ops in the same order: the difference is exactly zero."""
    r = subprocess.run([sys.executable, os.path.join(HERE, "pixel_pin_run.py")], capture_output=True, text=True, timeout=850)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["feature_gather"] == 0.0 and out["feature_fetch"] == 0.0
    assert [m["stage"] for m in out["multiscale"]] == [0, 1, 2, 3]
    for m in out["multiscale"]:
        assert m["max_abs"] == 0.0 and m["ref_max"] > 1.0 and m["nonzero"] > 0.2, m
    assert [m["channels"] for m in out["multiscale"]] == [32, 64, 128, 256]
