"""Drop-in for the reference's ``sdod.EfficientGN`` / ``efficient_group_norm`` (sdod/efficient_gn.py),
backed by the hand-written sm_100a GroupNorm kernels through ``torch.ops.sdod.group_norm``.

Interface kept from the reference (file:line = reference sdod/efficient_gn.py):
  * ``EfficientGN(num_groups, num_channels, eps=1e-5, affine=True, device=None, dtype=None, impl=None)`` (:34)
  * ValueError when ``num_channels % num_groups`` or ``impl`` is not one of None/'eff'/'ln'/'bn' (:37-40)
  * parameters ``weight``/``bias`` of shape [C], ones/zeros, state_dict-compatible with nn.GroupNorm (:47-59)
  * ``forward`` asserts ``shape[1] == num_channels`` and accepts any trailing spatial rank (:61-65)
  * functional ``efficient_group_norm(input, num_groups, weight=None, bias=None, eps=1e-5)`` (:29-30)
  * ONNX names ``sdod::GroupNorm`` / ``sdod::ParameterlessGroupNorm`` with attrs num_groups_i, eps_f (:14-26)

Only ``impl in (None, 'eff')`` compute GroupNorm in the reference; 'ln' drops the affine and 'bn' does not
normalise (:71-86) — those two are rejected here with NotImplementedError rather than
reproduced.  CUDA tensors only: there is no CPU fallback.
"""
import torch
import torch.nn as nn

from . import ops as _ops  # noqa: F401  (registers torch.ops.sdod.*)

ONNX_OP_GROUP_NORM = "sdod::GroupNorm"
ONNX_OP_PARAMETERLESS = "sdod::ParameterlessGroupNorm"


def efficient_group_norm(input, num_groups, weight=None, bias=None, eps=1e-5, silu=False, add_nc=None):
    """y = GroupNorm(input) [reference signature]; ``silu`` / ``add_nc`` expose the fused variants."""
    if (weight is None) != (bias is None):
        raise ValueError("weight and bias must both be given or both be None")
    if not input.is_cuda:
        from ._cabi import SdodError
        raise SdodError("sdod.EfficientGN runs on CUDA tensors only (hand-written sm_100a kernels; no CPU fallback)")
    return torch.ops.sdod.group_norm(input, num_groups, weight, bias, eps, silu, add_nc)


def onnx_symbolic(g, input, num_groups, weight, bias, eps):
    """Export-time mapping identical to EfficientGNFun.symbolic (efficient_gn.py:14-26)."""
    if weight is None:
        assert bias is None
        ret = g.op(ONNX_OP_PARAMETERLESS, input, num_groups_i=num_groups, eps_f=eps)
    else:
        assert bias is not None
        ret = g.op(ONNX_OP_GROUP_NORM, input, weight, bias, num_groups_i=num_groups, eps_f=eps)
    ret.setType(input.type())
    return ret


def register_onnx_symbolics(opset_version=13):
    """Map ``torch.ops.sdod.group_norm`` to the reference's ONNX nodes for ``torch.onnx.export(..., custom_opsets={"sdod": 1}, dynamo=False)``
    (reference tests/custom_export.py:21-30).  The fused extras (SiLU, temb add) have no node in the reference's op package
    (csrc/sdod_ops/config/group_norm.json) and refuse to export."""
    from torch.onnx import register_custom_op_symbolic, symbolic_helper

    @symbolic_helper.parse_args("v", "i", "v", "v", "f", "b", "v")
    def _group_norm(g, x, num_groups, weight, bias, eps, silu, add_nc):
        if silu or not symbolic_helper._is_none(add_nc):
            raise RuntimeError("sdod::group_norm: the fused SiLU / add_nc variants have no ONNX node in the reference op package")
        w_none, b_none = symbolic_helper._is_none(weight), symbolic_helper._is_none(bias)
        return onnx_symbolic(g, x, num_groups, None if w_none else weight, None if b_none else bias, eps)

    register_custom_op_symbolic("sdod::group_norm", _group_norm, opset_version)


class EfficientGN(nn.Module):
    def __init__(self, num_groups: int, num_channels: int, eps: float = 1e-5, affine: bool = True, device=None, dtype=None, impl=None) -> None:
        factory_kwargs = {"device": device, "dtype": dtype}
        super().__init__()
        if num_channels % num_groups != 0:
            raise ValueError("num_channels must be divisible by num_groups")
        if impl not in [None, "eff", "ln", "bn"]:
            raise ValueError('EfficientGN impl parameter should be one of None, "eff", "ln" or "bn"')
        self.num_groups = num_groups
        self.num_channels = num_channels
        self.eps = eps
        self.affine = affine
        self.impl = impl
        if self.affine:
            self.weight = nn.Parameter(torch.empty(num_channels, **factory_kwargs))
            self.bias = nn.Parameter(torch.empty(num_channels, **factory_kwargs))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self) -> None:
        if self.affine:
            nn.init.ones_(self.weight)
            nn.init.zeros_(self.bias)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        assert input.shape[1] == self.num_channels
        if self.impl in (None, "eff"):
            return efficient_group_norm(input, self.num_groups, self.weight, self.bias, self.eps)
        raise NotImplementedError(
            "impl=%r is not GroupNorm-equivalent in the reference (efficient_gn.py:71-86) and is not carried over" % self.impl)

    def extra_repr(self) -> str:
        return "{num_groups}, {num_channels}, eps={eps}, affine={affine}".format(**self.__dict__)
