"""Whole-model PARAMETER gradients at the north-star tolerance (1e-4), with the discrete selections pinned.

TomoSAR2Height is piecewise linear: given the pattern of ReLU masks, 2x2 max-pool winners and scatter-max argmax its
parameter gradients are smooth functions of the inputs, but the pattern itself flips under fp32 rounding (a
pre-activation of 1e-7 is positive in one implementation and negative in the other), and ONE flipped selection moves
single gradient entries by O(1e-3).  Round 1 bounded the whole-model gradients by 2e-3 and argued with such flips
without showing them.  This test separates the two effects:

  1. the CUDA path runs forward + backward while every selection it makes is recorded (the pre-activations are the
     operands the kernels receive, so the masks are exactly the ones the fused ReLU-on-load / mask epilogues apply);
  2. the CPU oracle is evaluated with those selections REPLAYED (oracle.Selections) -- same piecewise-linear branch;
  3. parameter gradients must then agree to 1e-4 of max |g| (the tolerance north_star states), for every parameter;
  4. the oracle's OWN selections are recorded too and the flips (entries where the two patterns differ) are counted
     and printed, together with the gradient deviation against the un-pinned oracle.
"""
import types

import pytest
import torch
import torch.nn.functional as F

import oracle
from cases import CASES, make_cfg, synthetic_cloud

pytestmark = pytest.mark.gpu

REL = 1e-4


class _Recorder:
    """Monkeypatches the call sites of the B200 model where a selection is made; tape entries in oracle layout."""

    def __init__(self):
        self.tape, self.topo, self._undo = [], None, []

    # -- layout helpers ------------------------------------------------------------------------------------
    def _rows_mask(self, x, x2=None):
        m = x > 0 if x2 is None else torch.cat([x > 0, x2 > 0], dim=-1)
        if m.dim() == 2 and self.topo is not None and m.shape[0] == self.topo.n_points:  # sorted rows -> (B, N, C)
            out = torch.empty_like(m)
            out[self.topo.perm.long()] = m
            return out.view(self.topo.B, self.topo.N, -1).cpu()
        return m.cpu()

    def _patch(self, mod, name, new):
        self._undo.append((mod, name, getattr(mod, name)))
        setattr(mod, name, new)

    def __enter__(self):
        import tomosar2height_b200.block.resnet as m_res
        import tomosar2height_b200.encoder.pointnet as m_pn
        import tomosar2height_b200.encoder.alto as m_alto
        import tomosar2height_b200.decoder.pixel as m_px
        import tomosar2height_b200.functional as T
        from tomosar2height_b200.linear import linear as real_linear, relu as real_relu
        from tomosar2height_b200.conv import apply_conv as real_conv
        from tomosar2height_b200.topology import Topology as RealTopology
        rec = self

        def linear(x, weight, bias=None, x2=None, relu_in=False, residual=None):
            if relu_in:
                rec.tape.append(("relu", rec._rows_mask(x.detach(), None if x2 is None else x2.detach())))
            return real_linear(x, weight, bias, x2=x2, relu_in=relu_in, residual=residual)

        def apply_conv(module, x, relu_in=False, residual=None):
            if relu_in:
                rec.tape.append(("relu", (x.detach() > 0).cpu()))
            return real_conv(module, x, relu_in=relu_in, residual=residual)

        def relu(x, *a, **k):
            rec.tape.append(("relu", (x.detach() > 0).cpu()))
            return F.relu(x, *a, **k)

        def leaky_relu(x, *a, **k):
            rec.tape.append(("relu", (x.detach() > 0).cpu()))
            return F.leaky_relu(x, *a, **k)

        f_proxy = types.SimpleNamespace(**{n: getattr(F, n) for n in dir(F) if not n.startswith("__")})
        f_proxy.relu, f_proxy.leaky_relu = relu, leaky_relu

        def topology(*a, **k):
            rec.topo = RealTopology(*a, **k)
            return rec.topo

        def seg_max_pool(rows, level, return_arg=False):
            pooled, arg = T.seg_max_pool(rows, level, return_arg=True)
            topo = rec.topo
            B, N, M = topo.B, topo.N, level.n_seg // topo.B
            a = arg.long().view(B, M, -1)
            base = (torch.arange(B, device=a.device) * N).view(B, 1, 1)
            a_pt = torch.where(a < 0, torch.full_like(a, N), topo.perm.long()[a.clamp(min=0)] - base)
            rec.tape.append(("argmax", a_pt.permute(0, 2, 1).contiguous().cpu()))
            return (pooled, arg) if return_arg else pooled

        t_proxy = types.SimpleNamespace(**{n: getattr(T, n) for n in dir(T) if not n.startswith("__")})
        t_proxy.seg_max_pool = seg_max_pool
        for mod in (m_res, m_pn, m_alto, m_px):
            self._patch(mod, "linear", linear)
        def relu_bounded(x, slope=0.0):
            rec.tape.append(("relu", (x.detach() > 0).cpu()))
            return real_relu(x, slope)

        for mod in (m_alto, m_px):
            self._patch(mod, "apply_conv", apply_conv)
            self._patch(mod, "F", f_proxy)
            self._patch(mod, "relu_bounded", relu_bounded)
        # ConvDecoder binds its activation at construction time: patched per instance in `attach`
        self._patch(m_pn, "Topology", topology)
        self._patch(m_pn, "T", t_proxy)
        return self

    def attach(self, model):
        """max-pool winners (forward hooks on the nn.MaxPool2d modules) and the decoder's bound activation"""
        rec = self

        def pool_hook(_m, inp, _out):
            _, idx = F.max_pool2d(inp[0].detach(), 2, 2, return_indices=True)
            rec.tape.append(("pool", idx.cpu()))

        self._hooks = [m.register_forward_hook(pool_hook) for m in model.modules() if isinstance(m, torch.nn.MaxPool2d)]
        for m in model.modules():
            if callable(getattr(m, "act", None)) and not isinstance(m.act, torch.nn.Module):
                def act(x, _inner=m.act):
                    rec.tape.append(("relu", (x.detach() > 0).cpu()))
                    return _inner(x)

                self._undo.append((m, "act", m.act))
                m.act = act

    def __exit__(self, *exc):
        for obj, name, old in reversed(self._undo):
            setattr(obj, name, old)
        for h in getattr(self, "_hooks", []):
            h.remove()
        return False


# wide enough that the 3xFP16 kernels run (K >= 128, n_out > 64 from the third ALTO level on) besides the 3xTF32 ones
WIDE = dict(cfg=dict(reso=32, depth=4, start=32, output_size=64), B=2, N=2500, seed=31)


@pytest.mark.parametrize("spec", [CASES["berlin_small"], WIDE], ids=["berlin_small", "wide"])
def test_parameter_gradients_match_at_pinned_selections(spec):
    import tomosar2height_b200 as t2h
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        cfg = make_cfg(**spec["cfg"])
        params = oracle.synth_state_dict(oracle.reference_param_shapes(cfg), seed=spec["seed"])
        model = t2h.TomoSAR2Height(cfg)
        model.load_state_dict(params)
        model = model.cuda()
        B, N = spec["B"], spec["N"]
        size = cfg.model.decoder_pixel_kwargs.output_size
        cloud = synthetic_cloud(B, N, seed=spec["seed"] + 7)
        g = torch.Generator().manual_seed(3)
        w = torch.randn(B, size, size, 1, generator=g)  # a smooth functional: the L1 loss adds sign() selections of its own

        with _Recorder() as rec:
            rec.attach(model)
            pa, _ = model(input_cloud=cloud.cuda())
            (pa * w.cuda()).mean().backward()
        gpu_tape = rec.tape

        def oracle_grads(selections):
            P = {k: v.clone().requires_grad_(True) for k, v in params.items()}
            out, _ = oracle.oracle_forward(P, cfg, cloud, selections=selections)
            (out * w).mean().backward()
            return out.detach(), {k: v.grad for k, v in P.items()}

        own = oracle.Selections("record")
        pa_own, g_own = oracle_grads(own)
        replay = oracle.Selections("replay", gpu_tape)
        pa_pin, g_pin = oracle_grads(replay)
        assert replay.pos == len(gpu_tape) == len(own.tape), "the CUDA path and the oracle make the same sequence of selections"

        flips = total = 0
        for (k1, a), (k2, b) in zip(gpu_tape, own.tape):
            assert k1 == k2 and a.shape == b.shape, (k1, k2, a.shape, b.shape)
            flips += int((a != b).sum())
            total += a.numel()
        scale = pa_own.abs().max()
        assert (pa.detach().cpu() - pa_own).abs().max() <= REL * scale      # heights: pinned or not
        assert (pa.detach().cpu() - pa_pin).abs().max() <= REL * scale

        worst_pin = worst_own = 0.0
        for name, p in model.named_parameters():
            if g_pin.get(name) is None or p.grad is None:
                continue
            denom = g_pin[name].abs().max().clamp(min=1e-30)
            e_pin = ((p.grad.cpu() - g_pin[name]).abs().max() / denom).item()
            e_own = ((p.grad.cpu() - g_own[name]).abs().max() / denom).item()
            worst_pin, worst_own = max(worst_pin, e_pin), max(worst_own, e_own)
            assert e_pin <= REL, f"{name}: {e_pin:.2e} at pinned selections"
        print(f"selections: {total}, flipped between the CUDA path and the CPU oracle: {flips}; worst parameter-gradient "
              f"deviation {worst_pin:.2e} at pinned selections vs {worst_own:.2e} against the oracle's own pattern")
        if flips == 0:
            assert worst_own <= REL
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
