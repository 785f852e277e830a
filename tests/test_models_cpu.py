"""CPU: the drop-in modules' plumbing (everything around the hot path) against the real reference's
golden outputs.  The four hot-path operators are routed to the oracle restatement FOR THIS TEST ONLY
(helpers.use_oracle_ops); the product itself has no CPU path."""
import copy
import hashlib

import numpy as np
import pytest
import torch

import helpers
from oracle.model_fixture import run_model


def _keys_digest(model):
    lines = ["%s %s" % (k, tuple(v.shape)) for k, v in model.state_dict().items()]
    return len(lines), hashlib.sha1("\n".join(lines).encode()).hexdigest()[:16]


def test_state_dict_contract():
    """key/shape lists of the reference modules (411 / 482 entries)."""
    assert _keys_digest(helpers.seeded_model("s")) == (411, "e91e8cbe371006d4")
    assert _keys_digest(helpers.seeded_model("lw")) == (482, "ea7a5f8397e36f1e")


def test_ofw_owns_the_reference_parameters():
    ofw = helpers.seeded_model("lw").OFW
    keys = sorted(ofw.state_dict())
    assert len(keys) == 22 and "flow_make.weight" in keys and "down.7.num_batches_tracked" in keys
    assert sum(p.numel() for p in ofw.parameters()) == 3168


@pytest.mark.parametrize("kind", ["s", "lw"])
def test_module_matches_reference_golden(kind, monkeypatch):
    helpers.use_oracle_ops(monkeypatch)
    torch.set_num_threads(8)
    model = helpers.seeded_model(kind)
    x1, x2 = helpers.seeded_pair(2)
    got = run_model(model, x1, x2, helpers.seeded_labels(2))
    want = helpers.load_golden("model_%s.npz" % kind)
    for k in want.files:
        a, b = want[k], got[k]
        if a.dtype == np.uint8:   # packed change maps: identical within 0.1 % of pixels (north_star)
            diff = np.unpackbits(a ^ b).sum()
            assert diff <= 0.001 * a.size * 8, (k, diff)
        else:
            err = np.abs(a.astype(np.float64) - b).max() / max(1e-12, np.abs(a).max())
            assert err < 2e-4, (k, err)


def test_against_live_reference_when_present(monkeypatch):
    ref = helpers.import_reference()
    if ref is None:
        pytest.skip("reference tree not present on this machine")
    helpers.use_oracle_ops(monkeypatch)
    mine = helpers.seeded_model("lw").eval()
    theirs = ref[1].SMOW_Net_LW()
    theirs.load_state_dict(copy.deepcopy(mine.state_dict()), strict=True)
    theirs.eval()
    x1, x2 = helpers.seeded_pair(1, seed=5)
    with torch.no_grad():
        assert (mine(x1, x2) - theirs(x1, x2)).abs().max() <= 1e-6
