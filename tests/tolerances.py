"""The tolerances of BASELINE.json's north_star, in one place.

* boxes / keypoints / landmarks: 1e-3 of the network input size;
* scores / flags: 1e-3 absolute ON THE SCORE.  Raw logits are compared at 4e-3 - the sigmoid's slope is at most 1/4, so
  4e-3 on a logit is at most 1e-3 on the score - plus 2e-5 of the logit's magnitude: saturated logits reach |v| ~ 270
  (score exactly 0 or 1), where f32 rounding of the reference engines themselves is of that order.
"""
import numpy as np

TOL = 1e-3
LOGIT_TOL = 4e-3
LOGIT_REL = 2e-5


def sigmoid(v):
    v = np.asarray(v, np.float64)
    return 1.0 / (1.0 + np.exp(-v))


def logit_excess(got, want, extra=0.0):
    """max over elements of |got - want| - (4e-3 + 2e-5 |want| + extra); <= 0 means within tolerance."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return float((np.abs(got - want) - (LOGIT_TOL + LOGIT_REL * np.abs(want) + extra)).max())


def assert_logits_close(got, want, extra=0.0, what=""):
    ex = logit_excess(got, want, extra)
    assert ex <= 0.0, (what, "logit error exceeds 4e-3 + 2e-5|v| by", ex)
    err = float(np.abs(sigmoid(got) - sigmoid(want)).max())
    assert err <= TOL + extra, (what, "score error", err)
