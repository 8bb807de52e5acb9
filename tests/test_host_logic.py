"""CPU tests of the host-side mirror of the reference interface (no kernels are launched)."""
import pytest
import torch

import inverse_flow_b200 as ifb
from inverse_flow_b200 import functional as IF
from inverse_flow_b200 import inv_conv_with_bp
from inverse_flow_b200.layers import Inv_FlowUnit, inv_flow_no_pad, inv_flow_with_pad


def test_default_groups_follows_the_reference_kernels():
    assert ifb.default_groups(4) == 4 and ifb.default_groups(12) == 4 and ifb.default_groups(48) == 4
    assert ifb.default_groups(1) == 1 and ifb.default_groups(3) == 1


def test_cpu_tensors_are_rejected_like_CHECK_CUDA():
    x = torch.randn(1, 4, 5, 5)
    w = torch.randn(4, 4, 3, 3)
    for fn in (lambda: IF.inverse(x, w), lambda: IF.conv(x, w), lambda: IF.bwd_input(x, w),
               lambda: inv_conv_with_bp.inverse(x, w, torch.zeros_like(x))):
        with pytest.raises(RuntimeError, match="CUDA"):
            fn()


def test_reference_module_surface():
    for name in ("inverse", "forward", "dy", "dw"):
        assert callable(getattr(inv_conv_with_bp, name))


@pytest.mark.parametrize("cls,kw", [(inv_flow_no_pad, {}), (inv_flow_with_pad, {"order": "TL"}),
                                    (inv_flow_with_pad, {"order": "BR"})])
def test_layer_parameters_match_the_reference_init(cls, kw):
    torch.manual_seed(0)
    layer = cls(8, 8, (3, 3), **kw)
    assert list(layer.state_dict().keys()) == ["weight_fwd"]
    w = layer.weight_fwd.data
    assert w.shape == (8, 8, 3, 3)
    assert torch.all(w[:, -1, -1, -1] == 1.0)                  # inv_conv.py:168-170
    # nn.init.dirac_ puts the 1 at the kernel's spatial centre (1,1) of a 3x3, i.e. on the
    # shift-(1,1) tap, not on the implicit diagonal (inv_conv.py:156-165)
    diag = torch.stack([w[c, c, 1, 1] for c in range(8)])
    assert torch.allclose(diag, torch.ones(8), atol=0.05)
    assert float(w[0, 1, 0, 0].abs()) < 0.05
    mask = layer.get_mask()
    assert mask[3, 3, -1, -1] == 0 and mask[3, 4, -1, -1] == 0 and mask[3, 2, -1, -1] == 1
    assert mask[3, 5, 0, 0] == 1
    layer.weight_fwd.grad = torch.ones_like(w)
    layer.reset_gradients()
    assert torch.equal(layer.weight_fwd.grad, mask)
    assert layer.logdet(torch.zeros(2, 8, 4, 4)) == 0.0
    assert layer.kernel_size == (3, 3)


def test_unit_requires_multiple_of_four_channels():
    with pytest.raises(AssertionError):
        Inv_FlowUnit(6, 6, (3, 3))
    unit = Inv_FlowUnit(8, 8, 3)
    assert [c.order for c in (unit.conv_tl, unit.conv_tr, unit.conv_bl, unit.conv_br)] == ["TL", "TR", "BL", "BR"]
    assert len(unit.state_dict()) == 4


def test_non_square_layer_is_refused():
    with pytest.raises(ValueError):
        inv_flow_no_pad(4, 8, (3, 3))


def test_if_glow_model_structure():
    from inverse_flow_b200 import glow
    model, shape, batch = glow.build("if_glow_mnist", coupling_width=16, groups=1)
    assert shape == (1, 28, 28) and batch == 100
    assert len(model.inv_layers) == 32
    assert [l.in_channels for l in model.inv_layers[:16]] == [4] * 16
    assert [l.in_channels for l in model.inv_layers[16:]] == [8] * 16      # split prior halves 16 -> 8... 4*4/2
    x = torch.randn(2, 8, 4, 4)
    sq = glow.Squeeze()
    y, _ = sq(x)
    assert y.shape == (2, 32, 2, 2) and torch.equal(sq.reverse(y), x)
    c = glow.Coupling(8, 16)
    yc, ld = c(torch.randn(2, 8, 4, 4))
    assert ld.shape == (2,)


def test_fincflow_layer_surface():
    from inverse_flow_b200.layers import Finc_FlowUnit, PaddedConv2d
    torch.manual_seed(0)
    m = PaddedConv2d(4, 4, (3, 3), order='BR')
    assert list(m.state_dict().keys()) == ["conv.weight"]
    tl = m.tl_weight()
    assert torch.all(torch.stack([tl[c, c, -1, -1] for c in range(4)]) == 1.0)      # unit diagonal
    assert all(torch.all(tl[c, c + 1:, -1, -1] == 0) for c in range(4))              # masked upper triangle
    assert torch.equal(torch.flip(m.conv.weight.data, [2, 3]), tl)                   # stored pre-flipped
    out, ld = m(torch.randn(2, 4, 5, 6))                                             # forward is plain torch
    assert out.shape == (2, 4, 5, 6) and ld == 0.0
    unit = Finc_FlowUnit(8, 8, 3)
    assert sorted(unit.state_dict().keys()) == ["conv_bl.conv.weight", "conv_br.conv.weight",
                                                "conv_tl.conv.weight", "conv_tr.conv.weight"]
    with pytest.raises(AssertionError):
        Finc_FlowUnit(6, 6, 3)


def test_squeeze_helpers_match_the_reference_formulas():
    """functional.space_to_depth / depth_to_space (the unfused path of the fused ActNorm/Squeeze layer and its tests)
    restate reference inf/layers/squeeze.py:5-24: checked against an index-by-index definition on the CPU"""
    import torch
    from inverse_flow_b200 import functional as IF
    x = torch.arange(2 * 3 * 4 * 6, dtype=torch.float32).view(2, 3, 4, 6)
    y = IF.space_to_depth(x)
    assert y.shape == (2, 12, 2, 3)
    for c in range(3):
        for i in range(2):
            for j in range(2):
                assert torch.equal(y[:, c * 4 + i * 2 + j], x[:, c, i::2, j::2])
    assert torch.equal(IF.depth_to_space(y), x)
