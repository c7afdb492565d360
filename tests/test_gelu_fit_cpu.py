"""CPU: the three-coefficient tanh form that the sm_100a epilogues evaluate for erf-GELU (csrc/a4r_common.cuh: kGeluC0..2)
against the exact definition 0.5 x (1 + erf(x / sqrt 2)) that the reference's nn.GELU() / ACT2FN['gelu'] compute
(Downstream/Text/model/encoders.py:45,57; modules.py:122-125)."""
import math
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _coefficients():
    src = open(os.path.join(ROOT, "adapter4rec_b200", "csrc", "a4r_common.cuh")).read()
    m = re.search(r"kGeluC0 = ([-0-9.e]+)f, kGeluC1 = ([-0-9.e]+)f, kGeluC2 = ([-0-9.e]+)f", src)
    assert m, "GELU coefficients not found in a4r_common.cuh"
    return [float(g) for g in m.groups()]


def test_tanh_form_matches_erf_gelu_and_its_derivative():
    c0, c1, c2 = _coefficients()
    x = np.linspace(-40.0, 40.0, 400001)
    erf = np.vectorize(math.erf)
    phi_exact = 0.5 * (1.0 + erf(x / math.sqrt(2.0)))
    x2 = np.minimum(x * x, 64.0)
    cdf = 0.5 + 0.5 * np.tanh(x * (c0 + x2 * (c1 + x2 * c2)))
    # forward: |gelu - gelu_exact| = |x| |dPhi| <= 3e-5 everywhere (bf16 half-ulp of the result is 2e-3 |gelu|)
    assert np.abs(x * (cdf - phi_exact)).max() <= 3e-5
    # backward: Phi + x phi with the exact density
    assert np.abs(cdf - phi_exact).max() <= 6e-5
    # saturation is monotone and exact far out (the clamp of x^2 keeps the negative x^4 coefficient from turning the
    # argument of tanh around)
    assert cdf[-1] == 1.0 and cdf[0] == 0.0 and np.all(np.diff(cdf) >= -1e-12)
