"""Process-wide flag namespace (the reference re-parses sys.argv in four modules at import;
here one Namespace is shared and can be replaced programmatically)."""
from .utility.parser import parse_args

_args = None


def get_args():
    global _args
    if _args is None:
        _args = parse_args([])
    return _args


def set_args(ns):
    global _args
    _args = ns
    return ns
