import logging as _pylogging


def is_ftfy_available():
    return True  # oracle/refshim/ftfy: identity on the ASCII prompts the fixtures use


def is_torch_xla_available():
    return False


class logging:  # noqa: N801 -- diffusers.utils.logging module surface
    @staticmethod
    def get_logger(name):
        return _pylogging.getLogger(name)


def replace_example_docstring(example_docstring):
    def deco(fn):
        return fn
    return deco
