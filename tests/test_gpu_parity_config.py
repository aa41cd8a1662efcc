"""Parity at the BENCHMARKED configuration and precision (VERDICT r1 "next" 1a): BASELINE config 2 exactly as bench.py
runs it -- 20 fragments x 1000 points, T = 100 DDPM steps, one verifier pass, replayed noise -- against the CPU
oracle (pinned bit for bit to the reference's own modules), in every precision mode of the engine.

Numbers measured on a B200 (profiles/r2_parity_config2.json, `python tests/parity_config.py`):

  mode   teacher-forced |d eps| (max / median)   VQ codes equal   steps with a flip   free-running final pose error
  fp32   2.0e-7 / 1.3e-7                         100 %            0 of 100            6.8e-5
  tc32   5.2e-4 / 1.2e-6                         99.990 %         18 of 100           1.3e-4
  bf16   2.2e-3 / 1.2e-3                         96.56 %          100 of 100          8.4e-3

fp32 (SIMT) meets the north star's 1e-4 outright.  tc32 (tensor cores, bf16 hi/lo split operands, 2^-16 per product)
sits at the flip-limited floor: with identical inputs its eps differs from the oracle's by ~1e-6 except on the 18
steps where one of the 2000 VQ code searches lands on the other side of a near-tie (then up to 5e-4), and the DDPM
recursion amplifies the fp32 mode's own 2e-7 per-step differences to 7e-5 over 100 steps.  bf16 is the fast mode.
The verifier's accept decisions are identical in every mode."""
import json
import os

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

# (teacher-forced eps max, eps median, min VQ code match, free-running final pose error)
TOL = {"fp32": (1e-5, 1e-6, 0.9999, 1e-4),
       "tc32": (2e-3, 1e-5, 0.9995, 5e-4),
       "bf16": (1e-2, 5e-3, 0.95, 3e-2)}


@pytest.fixture(scope="module")
def config2_report():
    from parity_config import report
    lines = []
    r = report(("fp32", "tc32", "bf16"), log=lines.append)
    print("\n" + "\n".join(lines))
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        json.dump(r, open(os.path.join(out, "parity_config2_from_test.json"), "w"), indent=1)
    return r


@pytest.mark.parametrize("mode", ["fp32", "tc32", "bf16"])
def test_config2_as_benchmarked_vs_oracle(config2_report, mode):
    tf, fr = config2_report[mode]["teacher_forced"], config2_report[mode]["free_running"]
    eps_max, eps_med, code, pose = TOL[mode]
    assert tf["eps_max"] <= eps_max and tf["eps_median"] <= eps_med, tf
    assert tf["code_match"] >= code, tf
    assert tf["fps_centroid_match"] == 1.0, tf          # FPS / ball-query indices are bit-exact in every mode
    assert fr["pose_final"] <= pose and fr["pose_max_over_steps"] <= 2 * pose, fr
    assert fr["decisions_equal"] and fr["feature_max"] <= 1e-6, fr
    assert fr["logit_max"] <= 2e-4, fr
