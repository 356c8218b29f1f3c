"""Type names only: the harness hands the pipelines oracle-backed objects."""


class _Named:
    pass


class AutoencoderKLWan(_Named):
    pass


class WanTransformer3DModel(_Named):
    pass


class AutoencoderKLCogVideoX(_Named):
    pass


class CogVideoXTransformer3DModel(_Named):
    pass


class AutoencoderKLHunyuanVideo(_Named):
    pass


class HunyuanVideoTransformer3DModel(_Named):
    pass
