"""Stand-in for ``ftfy`` (absent offline; wan:96-99 calls ``ftfy.fix_text``).  The fixtures use plain-ASCII prompts, which
ftfy leaves unchanged."""


def fix_text(text):
    return text
