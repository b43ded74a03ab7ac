"""
GPU: the reference's q-SVGP comparisons (reference tests/models/test_tsvgp.py:45-88,123-131,168-188; CPU restatement and
rationale in tests/test_reference_svgp_pins.py) run through the CUDA path: the device t-SVGP after 20 natural-gradient steps at
lr = 1 must predict like GPflow's SVGP after 20 NaturalGradient(gamma=1) steps (oracle/svgp_oracle.py), and its M-step gradients
w.r.t. the kernel hyperparameters must equal the SVGP's, at the reference's own tolerance (decimal=4; cond(Kuu) = 2.6e16 here).
"""
import numpy as np
import pytest

from oracle import tsvgp_oracle as orc
from tests.test_reference_svgp_pins import LENGTH_SCALE, VARIANCE, _softplus_jacobian, _svgp_loss_grads, _tsvgp_qsvgp_optim_setup

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("binary_labels", [False, True])
@pytest.mark.parametrize("num_latent_gps", [1, 2])
def test_device_tsvgp_matches_qsvgp_after_20_natgrad_steps(num_latent_gps, binary_labels):
    import tsvgp_b200 as tb
    ref_t, qsvgp, (X, Y) = _tsvgp_qsvgp_optim_setup(num_latent_gps, binary_labels)
    dev = tb.t_SVGP(ref_t.kernel, ref_t.likelihood, orc.InducingPoints(X.copy()), num_latent_gps=num_latent_gps)
    for _ in range(20):
        dev.natgrad_step((X, Y), lr=1.0)
    mu_d, var_d = dev.predict_f(X + 0.1)
    mu_q, var_q = qsvgp.predict_f(X + 0.1)
    np.testing.assert_array_almost_equal(mu_d, mu_q, decimal=4)
    np.testing.assert_array_almost_equal(var_d, var_q, decimal=4)
    np.testing.assert_almost_equal(dev.elbo((X, Y)), qsvgp.elbo((X, Y)), decimal=4)
    # test_tsvgp.py:168-188 : hyperparameter gradients of -ELBO, w.r.t. the unconstrained (softplus) variables
    _, g = dev.elbo_and_grad((X, Y))
    grads_d = -np.array([g["variance"] * _softplus_jacobian(VARIANCE), float(np.sum(g["lengthscales"])) * _softplus_jacobian(LENGTH_SCALE)])
    np.testing.assert_array_almost_equal(_svgp_loss_grads(qsvgp, (X, Y)), grads_d, decimal=4)
    dev.close()
