"""Mean-teacher EMA of the student parameters.

Mirror of SimiTeacherHook (reference: mmdet3d/core/utils/simi_teacher_hook.py:39-64,86-92):
  ema <- ema * (1 - m) + m * param,   m = min(momentum, (1 + step) / (warm_up + step))
and the student<->teacher parameter swap.  The reference walks ~220 tensors with two tiny
launches each (and 3 copies per tensor per swap, twice per step); here all parameters and all
EMA copies live in two flat fp32 buffers, so the update is ONE launch (nesie_ema_update) and the
swap is one 3-way flat copy.  Parameters only: BN running statistics are shared between student
and teacher, as in the reference.
"""
import torch

from . import _lib


ALIGN = 64   # floats


class TeacherEMA:

    def __init__(self, model, momentum=0.001, interval=1, warm_up=10, flat=None):
        """flat: a FlatGradDDP built with flatten_parameters=True; its flat parameter buffer (and layout)
        is then used as is instead of re-homing the parameters a second time."""
        assert isinstance(interval, int) and interval > 0
        assert 0 < momentum < 1
        self.momentum = momentum ** interval
        self.interval = interval
        self.warm_up = warm_up
        self.params = [p for _, p in model.named_parameters(recurse=True)]
        self.names = [n for n, _ in model.named_parameters(recurse=True)]
        dev = self.params[0].device
        if flat is not None:
            off = flat.offsets()
            self.offsets = [off[id(p)] for p in self.params]
            self.flat_param = flat.flat_params
            self.flat_ema = self.flat_param.clone()
            return
        # every parameter starts on a 256-byte boundary of the flat buffer: the kernels read weights and
        # BatchNorm vectors with 16-byte accesses / TMA (the padding floats are zero in both buffers)
        self.offsets, total = [], 0
        for p in self.params:
            self.offsets.append(total)
            total += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        _lib.need_cuda(self.params[0])
        for n, p in zip(self.names, self.params):
            assert p.dtype == torch.float32 and p.device == dev and p.layout == torch.strided, \
                f"TeacherEMA needs dense fp32 parameters on one CUDA device ({n})"
        # re-home every parameter into one flat buffer (views keep the module API unchanged)
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=dev)
        for p, off in zip(self.params, self.offsets):
            n = p.numel()
            self.flat_param[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + n].view_as(p.data)
        self.flat_ema = self.flat_param.clone()

    def ema_state_dict(self):
        """EMA copies under the reference's buffer names `ema_<param name with dots -> _>`
        (simi_teacher_hook.py:47-51)."""
        out = {}
        for name, p, off in zip(self.names, self.params, self.offsets):
            out[f"ema_{name.replace('.', '_')}"] = self.flat_ema[off:off + p.numel()].view_as(p.data)
        return out

    def ema_named_views(self):
        """{parameter name: EMA copy (a view of the flat buffer, shaped like the parameter)}: what
        torch.func.functional_call needs to run the module as the teacher without swapping."""
        return {name: self.flat_ema[off:off + p.numel()].view_as(p.data)
                for name, p, off in zip(self.names, self.params, self.offsets)}

    def load_ema_state_dict(self, state, strict=True):
        """Fill the EMA copies from `ema_*` entries of a reference checkpoint (the hook registers them
        as model buffers, simi_teacher_hook.py:47-51, so `epoch_N.pth` / `epoch_N_ema.pth` carry
        them).  Returns the consumed keys: drop them before a strict module.load_state_dict()."""
        used = []
        for key, dst in self.ema_state_dict().items():
            if key in state:
                dst.copy_(state[key].to(dst.device, dst.dtype).view_as(dst))
                used.append(key)
            elif strict:
                raise KeyError(f"missing EMA buffer {key}")
        return used

    def after_train_iter(self, curr_step):
        momentum = min(self.momentum, (1 + curr_step) / (self.warm_up + curr_step))
        if curr_step % self.interval != 0:
            return
        with torch.cuda.device(self.flat_param.device):
            _lib.call("nesie_ema_update", self.flat_param.numel(), _lib.ptr(self.flat_ema),
                      _lib.ptr(self.flat_param), float(1 - momentum), float(momentum),
                      _lib.stream())

    def swap(self):
        """Swap student and teacher weights in place (switch_to_teacher / switch_to_student)."""
        tmp = self.flat_param.clone()
        self.flat_param.copy_(self.flat_ema)
        self.flat_ema.copy_(tmp)
