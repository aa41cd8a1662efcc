"""Parity at the BENCHMARKED configuration and precision (VERDICT r1 "next" 1a): BASELINE config 2 exactly as bench.py
runs it -- 20 fragments x 1000 points, T = 100 DDPM steps, one verifier pass, replayed noise -- against the CPU
oracle (pinned bit for bit to the reference's own modules), in every precision mode of the engine.

Two views (numbers measured on a B200: profiles/r2_parity_config2.json, `python tests/parity_config.py`):

* TEACHER-FORCED (the engine is fed the oracle's x_t at every step; isolates one step's error) -- asserted tightly:

    mode   |d eps| max / median    VQ codes equal   steps with a flip
    fp32   1.9e-7 / 1.3e-7         100 %            0 of 100
    tc32   5.2e-4 / 2.2e-6         99.990 %         18 of 100
    bf16   2.2e-3 / 1.2e-3         96.56 %          100 of 100

* FREE-RUNNING (the engine runs all 100 steps on its own poses).  Until the first VQ-code flip the pose difference
  is a rounding error (<= 3e-7 in fp32 mode, <= 1e-5 in tc32); at a flip -- one of the 2000 code searches of a step
  landing on the other side of a near-tie, which a difference of 1e-7 in z_e is enough for -- the run becomes a
  different, equally valid trajectory and the difference jumps to ~1e-3, where it stays.  Whether a flip happens is
  luck: the same build measured 6.8e-5 (no flip) and 1.0e-3 (flip at step 2) in fp32 mode on two summation orders of
  the attention kernel.  The reference itself has this property between any two GPUs / library versions.  Asserted:
  pose error <= 1e-4 up to the first flip, flip-limited bound afterwards, identical verifier decisions."""
import json
import os

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

# (teacher-forced eps max, eps median, min VQ code match, free-running pose error before the first flip, after it)
TOL = {"fp32": (1e-5, 1e-6, 0.9999, 1e-5, 5e-3),
       "tc32": (2e-3, 1e-5, 0.9995, 1e-4, 5e-3),
       "bf16": (1e-2, 5e-3, 0.95, 3e-2, 3e-2)}


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
    eps_max, eps_med, code, pose_before, pose_after = TOL[mode]
    assert tf["eps_max"] <= eps_max and tf["eps_median"] <= eps_med, tf
    assert tf["code_match"] >= code, tf
    assert tf["fps_centroid_match"] == 1.0, tf          # FPS / ball-query indices are bit-exact in every mode
    assert fr["pose_max_before_first_flip"] <= pose_before, fr
    assert fr["pose_max_over_steps"] <= pose_after, fr
    if fr["first_code_flip_step"] < 0:                  # no discrete flip: the whole run is within rounding error
        assert fr["pose_final"] <= max(pose_before, 1e-4), fr
    assert fr["decisions_equal"] and fr["feature_max"] <= 1e-6, fr
    assert fr["logit_max"] <= 2e-4, fr


def test_config3_full_length_loop_vs_gpu_oracle():
    """BASELINE config 3's loop at full length (T = 100 DDPM steps per outer iteration, up to 6 iterations with verify /
    promote / merge and early exits; 4 objects of 20 / 16 / 12 / 9 fragments in ONE packed batch) against the oracle run
    object by object as eager fp32 PyTorch on the GPU (TF32 off), fp32 and tc32 modes.

    Asserted: the agglomeration decisions -- outer iterations run, reference promotions, merge pivots -- are identical
    for every object, and the final poses of the objects that did not merge agree within the flip-limited bound of
    the config-2 test (measured 6e-5 .. 4e-4 in fp32 mode, 8e-5 .. 1.9e-3 in tc32).  An object that merged continues
    on a re-sampled cloud whose random-start FPS begins at int(u * M), M = points kept by the intersection filter
    (node_merge_utils.py:216-220): a pose difference of 1e-4 can move one point across the filter's 1e-3 threshold,
    M changes by one, the start index changes, and the 1000-point re-sample -- and every pose after it -- is a different,
    equally valid one (measured: O(1) pose difference with identical decisions; profiles/r2_parity_loop.json).  The
    merge stage itself is pinned teacher-forced (tests/test_gpu_kernels.py::test_merge_filter_matches_oracle,
    tests/test_gpu_engine.py::test_loop_full_batch_with_merges_vs_oracle)."""
    from parity_config import loop_report
    lines = []
    r = loop_report(("fp32", "tc32"), log=lines.append)
    print("\n" + "\n".join(lines))
    for mode in ("fp32", "tc32"):
        rows = r[mode]
        assert all(row["decisions_equal"] for row in rows), (mode, rows)
        assert any(row["iterations"] > 2 for row in rows) and any(row["merged"] > 0 for row in rows), rows
        for row in rows:
            if row["merged"] == 0:
                assert row["pose_err"] <= 5e-3 and row["trajectory_err"] <= 5e-3, (mode, row)
