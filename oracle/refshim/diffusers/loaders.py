class WanLoraLoaderMixin:
    pass


class CogVideoXLoraLoaderMixin:
    pass


class HunyuanVideoLoraLoaderMixin:
    pass
