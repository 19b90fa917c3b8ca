"""Invalidation of the folded-weight launch plans (fusion_engine.py / adapnet_engine.py) and of CUDA graphs captured
over them.  A plan bakes the parameters of a module at build time, so it must die whenever they can have changed:

  * train() / .to() / .cuda() / load_state_dict() called on the module itself (the overrides in the owners);
  * load_state_dict() called on a PARENT -- nn.Module.load_state_dict recurses through _load_from_state_dict and never
    calls the child's override, but it does run every sub-module's load_state_dict post-hooks
    (the reference does exactly this on resume: pipeline.load_state_dict(checkpoint), train_fusion.py:114);
  * in-place updates (optimizer.step(), p.data.copy_()) while the module stays in eval mode: every tensor carries a
    version counter that such writes bump, so the owner compares a fingerprint of (data_ptr, _version) sums.
"""


class EngineOwner:
    """Mixin for an nn.Module that owns launch plans.  Subclasses implement _drop_engines()."""
    _fp_tensors = None
    _fp_value = None
    _fp_hooked = False
    precision = 'parity'        # 'parity': 3xTF32, ~1e-6 of fp32 (default); 'fast': 1xTF32, ~1e-3 (BASELINE.json configs[2])

    def set_precision(self, mode):
        """Arithmetic of the tensor-core convolutions of this module's launch plans.  'parity' (default) is the 3xTF32
        split product (fp32-grade, meets the 1e-4 tolerance on logits / TSDF updates); 'fast' issues one TF32 MMA per MAC
        (3x less tensor work, ~1e-3 relative) and is judged on label agreement / mIoU / F1 instead."""
        if mode not in ('parity', 'fast'):
            raise ValueError("precision must be 'parity' or 'fast'")
        if mode != self.precision:
            self.precision = mode
            self._invalidate()
        return self

    @property
    def conv_flags(self):
        return 64 if self.precision == 'fast' else 0

    def _drop_engines(self):            # pragma: no cover - overridden
        raise NotImplementedError

    def _invalidate(self):
        self._fp_tensors = None
        self._fp_value = None
        self._drop_engines()

    def _hook_load_state_dict(self):
        if not self._fp_hooked:
            self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())
            self._fp_hooked = True

    def params_fingerprint(self):
        """Cheap identity + version digest of every parameter and buffer (tens of microseconds for ~1000 tensors)."""
        ts = self._fp_tensors
        if ts is None:
            ts = self._fp_tensors = [t for t in list(self.parameters()) + list(self.buffers())]
        a = v = 0
        for t in ts:
            a += t.data_ptr()
            v += t._version
        return (len(ts), a, v)

    def engines_current(self):
        """Drop the plans if a parameter was replaced or written since they were built; True if they survived."""
        fp = self.params_fingerprint()
        if self._fp_value is None:
            self._fp_value = fp
            return True
        if fp != self._fp_value:
            self._drop_engines()
            self._fp_tensors = None
            self._fp_value = self.params_fingerprint()
            return False
        return True
