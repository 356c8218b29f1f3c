import contextlib

import torch


class _Bar:
    def update(self, n=1):
        pass


class DiffusionPipeline:
    """register_modules / _execution_device / progress_bar / maybe_free_model_hooks: all the loops touch."""

    def register_modules(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def _execution_device(self):
        return torch.device(getattr(self, "_device", "cpu"))

    @contextlib.contextmanager
    def progress_bar(self, iterable=None, total=None):
        yield _Bar()

    def maybe_free_model_hooks(self):
        pass
