"""Our kernels against the reference's OWN CUDA extension, both running on the same GPU.

oracle/build_ref_cuda.py compiles inf/utils/inv_conv_cuda/inv_conv_with_bp_{general.cpp,
kernel_general.cu} for sm_100a into the git-ignored oracle/_ref/ (it travels to the GPU box
as a built file).  The literal kernels hard-code 4 channel groups (kernel_general.cu:94-98) and
are self-consistent for C == 4: there the reference's `inverse` and
`forward` ARE the math contract with groups=4, so our kernels must agree with the reference
binary to fp32 rounding.  The literal `dy` computes L^-1 g, i.e. it must
equal our inverse applied to g.  For C == 8 the literal inverse reads its own channel
(kernel_general.cu:61); that mode exists only in the oracle (literal_inverse), which these
tests pin to the reference binary.  Square images only: the reference's diagonal indexing
assumes H == W (kernel_general.cu:41-48).
"""
import pytest
import torch

from oracle import oracle, ref_cuda
import inverse_flow_b200.functional as F

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_cuda.available(), reason="oracle/_ref/ has no compiled reference extension")]

TOL = 1e-5      # max relative error, the north-star fp32 bar


def _weight(C, k, seed, scale=0.2):
    g = torch.Generator().manual_seed(seed)
    w = (torch.rand(C, C, k, k, generator=g) - 0.5) * scale
    return w


def _rel(a, b):
    return oracle.max_rel_err(a.detach().cpu().numpy(), b.detach().cpu().numpy())


CASES = [(3, 4, 5, 3), (100, 4, 14, 2), (16, 4, 16, 3), (7, 4, 9, 4), (2, 4, 32, 3)]


@pytest.mark.parametrize("B,C,H,k", CASES)
def test_inverse_matches_reference_binary(B, C, H, k):
    ref = ref_cuda.load()
    torch.manual_seed(B * 131 + H)
    x = torch.randn(B, C, H, H, device="cuda")
    w = _weight(C, k, 5 + k).cuda()
    y_ref = ref.inverse(x, w, torch.zeros_like(x))[0]
    y = F.inverse(x, w, groups=4)
    torch.cuda.synchronize()
    assert _rel(y, y_ref) <= TOL
    # and against the CPU oracle, which the golden vectors pin to the reference's CPU solvers
    y_cpu = oracle.inverse(x.cpu().numpy(), w.cpu().numpy(), groups=4)
    assert oracle.max_rel_err(y_ref.cpu().numpy(), y_cpu) <= TOL


@pytest.mark.parametrize("B,C,H,k", CASES)
def test_conv_matches_reference_binary(B, C, H, k):
    ref = ref_cuda.load()
    torch.manual_seed(B * 17 + H)
    y = torch.randn(B, C, H, H, device="cuda")
    w = _weight(C, k, 9 + k).cuda()
    x_ref = ref.forward(y, w, torch.zeros_like(y))[0]
    x = F.conv(y, w, groups=4)
    torch.cuda.synchronize()
    assert _rel(x, x_ref) <= TOL


@pytest.mark.parametrize("B,C,H,k", [(3, 4, 5, 3), (20, 4, 14, 2), (4, 4, 12, 3)])
def test_literal_dy_is_inverse_of_g(B, C, H, k):
    ref = ref_cuda.load()
    torch.manual_seed(B + H)
    g = torch.randn(B, C, H, H, device="cuda")
    w = _weight(C, k, 21 + k).cuda()
    dy_ref = ref.dy(g, w, torch.zeros_like(g), torch.zeros_like(g))[0]
    ours = F.inverse(g, w, groups=4)
    torch.cuda.synchronize()
    assert _rel(ours, dy_ref) <= 5e-5           # the literal path sums (HW) products per pixel in fp32
    lit = oracle.literal_dy(g.cpu().numpy(), w.cpu().numpy())
    assert oracle.max_rel_err(dy_ref.cpu().numpy(), lit) <= 5e-5


@pytest.mark.parametrize("B,C,H,k", [(2, 8, 6, 3), (5, 12, 8, 2)])
def test_oracle_literal_mode_pinned_to_reference_binary(B, C, H, k):
    ref = ref_cuda.load()
    torch.manual_seed(C + H)
    x = torch.randn(B, C, H, H, device="cuda")
    w = _weight(C, k, 33 + k, scale=0.1).cuda()
    y_ref = ref.inverse(x, w, torch.zeros_like(x))[0]
    torch.cuda.synchronize()
    lit = oracle.literal_inverse(x.cpu().numpy(), w.cpu().numpy())
    assert oracle.max_rel_err(y_ref.cpu().numpy(), lit) <= TOL


def test_reference_round_trip_through_our_kernels():
    """reference inverse -> our conv and our inverse -> reference conv both give x back (C = 4)"""
    ref = ref_cuda.load()
    torch.manual_seed(3)
    x = torch.randn(8, 4, 10, 10, device="cuda")
    w = _weight(4, 3, 77).cuda()
    y_ref = ref.inverse(x, w, torch.zeros_like(x))[0]
    assert _rel(F.conv(y_ref.contiguous(), w, groups=4), x) <= TOL
    y = F.inverse(x, w, groups=4)
    assert _rel(ref.forward(y, w, torch.zeros_like(y))[0], x) <= TOL
