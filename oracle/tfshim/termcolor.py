"""termcolor stand-in for the reference's src/core/out.py (not installed in this image).  TEST INFRASTRUCTURE ONLY."""


def colored(text, *args, **kw_args):
    return text
