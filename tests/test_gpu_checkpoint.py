"""GPU: an ensemble loaded from a PE.save-style checkpoint predicts exactly like the same weights
uploaded directly (cmbpo_b200.load_pe)."""
import numpy as np
import pytest

from oracle import cmbpo_oracle as orc
from test_checkpoint_cpu import write_like_pe_save

pytestmark = pytest.mark.gpu


def test_loaded_checkpoint_predicts_like_direct_upload(engine, tmp_path):
    import cmbpo_b200 as cb
    from cmbpo_b200 import _lib as L
    dyn, actor, v, vc = orc.make_problem(11, 17, 6, hidden=(64, 64))
    obs, act = orc.make_states(12, 300, 17, 6, dyn)
    x = np.concatenate([obs, act], -1)
    direct = cb.B200PE.from_arrays(engine, L.NET_DYN, dyn)
    m0, v0 = direct.predict_ensemble(x)
    write_like_pe_save(str(tmp_path), "BNN", 7, dyn, ("in", "out"), nll=True)
    loaded = cb.load_pe(engine, L.NET_DYN, str(tmp_path), "BNN", 7, elite_inds=dyn.elite_inds)
    m1, v1 = loaded.predict_ensemble(x)
    assert loaded.is_probabilistic and loaded.elite_inds == [int(i) for i in dyn.elite_inds]
    assert np.array_equal(m0, m1) and np.array_equal(v0, v1)
    # and it matches the oracle's forward within the fp32 tolerance of the other tests
    mo, vo = orc.pe_forward(dyn, x)
    assert np.allclose(m1, mo, rtol=1e-3, atol=1e-4) and np.allclose(v1, vo, rtol=1e-3, atol=1e-6)


def test_policy_loaded_from_savedmodel_variables_acts_like_direct_upload(engine, tmp_path):
    """load_policy (policies/cpo_policy.py:890-894 + utilities/logx.py:202-259 on the reading side): the actor and the
    value ensembles read from a SavedModel's variables bundle give bit-identical actions / values to the same arrays
    uploaded directly, and match the oracle."""
    import os
    import cmbpo_b200 as cb
    from cmbpo_b200 import tf_bundle as tb
    from test_tf_bundle_cpu import _policy_variables
    tens, actor, v, vc = _policy_variables(5)
    export = str(tmp_path / "policy")
    tb.write_bundle(os.path.join(export, "variables", "variables"), tens)
    obs, _ = orc.make_states(6, 257, 17, 6, None)
    eps = np.random.default_rng(7).standard_normal((257, 6)).astype(np.float32)
    direct = cb.B200Policy(engine)
    direct.load_actor(actor.W, actor.b, actor.log_std)
    direct.load_values(v, vc)
    a0 = direct.get_action_outs(obs, eps=eps)
    v0, vc0 = direct.get_v(obs), direct.get_vc(obs)
    loaded = cb.load_policy(engine, export, vf_activation="swish", vf_elites=v.elite_inds)
    a1 = loaded.get_action_outs(obs, eps=eps)
    for k in ("pi", "logp_pi", "v", "vc"):
        assert np.array_equal(np.asarray(a0[k]), np.asarray(a1[k])), k
    assert np.array_equal(v0, loaded.get_v(obs)) and np.array_equal(vc0, loaded.get_vc(obs))
    mu = orc.actor_mu(actor, obs)
    assert np.allclose(np.asarray(a1["pi_info"]["mu"]), mu, rtol=1e-3, atol=1e-4)
