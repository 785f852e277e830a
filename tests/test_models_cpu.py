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


def test_cyclic_frame_mix_composed_equals_oracle_matrix_form():
    """Row N4 on the CPU: the module-level composition (slices, 1x1x1 convolutions, adds, concat) that the CUDA kernels
    are tested against equals the oracle's matrix restatement of the reference (oracle/torch_ref.py), for both the
    Conv3d (SMOW_Net_LW conv_block_2_3d) and the biased ConvTranspose3d (SMOW_Net conv_trans_block_3d) flavours."""
    import torch
    from oracle import torch_ref
    from smow_net_b200.models import blocks
    torch.manual_seed(0)
    for make, cin, cout in ((lambda: torch.nn.Conv3d(12, 12, 1, bias=False), 12, 12),
                            (lambda: torch.nn.ConvTranspose3d(10, 6, 1, bias=True), 10, 6)):
        mods = [make().double() for _ in range(5)]
        x = torch.randn(2, cin, 4, 5, 7, dtype=torch.float64)
        got = blocks.cyclic_frame_mix(x, mods[4], mods[:4])
        bias = None
        if mods[4].bias is not None:
            bias = torch.stack([mods[4].bias + mods[(j + 1) % 4].bias for j in range(4)])
        want = torch_ref.ref_cyclic_frame_mix(x, blocks._mix_matrix(mods[4]), torch.stack([blocks._mix_matrix(m) for m in mods[:4]]), bias)
        assert float((got - want).abs().max()) <= 1e-12


def test_decoder_blocks_equal_the_reference_blocks():
    """With the same parameters (time convolutions perturbed away from their identity / zero initialisation) our
    TemporalDeconvMix / SpatialConvMix reproduce the reference's conv_trans_block_3d / conv_block_2_3d on the CPU."""
    import torch
    ref = helpers.import_reference()
    if ref is None:
        pytest.skip("reference tree not available")
    from smow_net_b200.models import blocks
    ref_lw = ref[1]
    torch.manual_seed(1)
    for ours, theirs, cin in ((blocks.TemporalDeconvMix(8, 8, wide=False), ref_lw.conv_trans_block_3d(8, 8), 8),
                              (blocks.SpatialConvMix(12, 8), ref_lw.conv_block_2_3d(12, 8), 12)):
        with torch.no_grad():
            for p in theirs.parameters():
                p.add_(torch.randn_like(p) * 0.1)
        ours.load_state_dict(theirs.state_dict(), strict=True)
        x = torch.randn(2, cin, 4, 6, 5)
        ours.train(); theirs.train()
        assert float((ours(x) - theirs(x)).abs().max()) <= 1e-5


def test_tokenizer_oracle_and_module_equal_the_reference_token_encoder():
    """Row N2 on the CPU: the oracle's tokenizer restatement (oracle/torch_ref.py::ref_semantic_tokens, what the CUDA
    kernels are tested against) plugged into the reference's own Transformer_Encoder reproduces its output, and so
    does our module with the same state_dict."""
    import torch
    ref = helpers.import_reference()
    if ref is None:
        pytest.skip("reference tree not available")
    from oracle import torch_ref
    from smow_net_b200.models import tokens
    torch.manual_seed(2)
    theirs = ref[1].Transformer_Encoder(in_chan=8).eval()
    ours = tokens.Transformer_Encoder(in_chan=8).eval()
    ours.load_state_dict(theirs.state_dict(), strict=True)
    x = torch.randn(2, 8, 4, 9, 11)
    with torch.no_grad():
        want = theirs(x)
        tok = torch_ref.ref_semantic_tokens(x, theirs.conv_a.weight, theirs.conv_a.bias) + theirs.pos_embedding.unsqueeze(0)
        via_oracle = theirs.transformer(tok.permute(0, 2, 1, 3).reshape(2, 8, 32))
        got = ours(x)
    assert float((via_oracle - want).abs().max()) <= 1e-6
    assert float((got - want).abs().max()) <= 1e-6
