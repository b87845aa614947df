"""Inert stand-in for the ``dgl`` package (oracle only; see oracle/__init__.py).

The reference's ``src/model.py:3`` does ``from dgl import function as fn`` and
only ever passes ``fn.copy_src / fn.mean / fn.max`` descriptors back into
``graph.pull`` -- they never compute anything themselves.
"""
from . import function  # noqa: F401
