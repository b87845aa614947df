"""Descriptors returned by the builtin message/reduce constructors."""
from collections import namedtuple

CopySrc = namedtuple("CopySrc", "src_field out_field")
Reduce = namedtuple("Reduce", "kind msg_field out_field")


def copy_src(src, out):
    return CopySrc(src, out)


copy_u = copy_src


def mean(msg, out):
    return Reduce("mean", msg, out)


def max(msg, out):  # noqa: A001 - mirrors dgl.function.max
    return Reduce("max", msg, out)


def sum(msg, out):  # noqa: A001
    return Reduce("sum", msg, out)
