"""Minimal stand-in for the DGL heterograph surface used by the reference.

Oracle / test infrastructure only.  Restates DGL's ``pull`` semantics as listed
in SURVEY.md Appendix A so that the reference's *own* ``PathConv.forward`` and
its UDFs (``model.py:88-116,138-213``) run unmodified on CPU:

* ``pull(v, mfunc, rfunc, apply_node_func, etype)`` touches only in-edges of
  ``v`` of edge type ``etype`` (reference call sites ``model.py:186-204``);
* builtin ``mean`` = sum / in-degree (zero in-degree -> 0), builtin ``max``
  with zero in-degree -> 0;
* a UDF reduce is degree-bucketed, mailbox shape ``(n, deg, D)``, rows with
  zero in-degree get 0;
* the apply function sees the node's current ndata plus the reduced field and
  only the dict it returns is written back, out of place (``index_copy``);
* an empty node list is a no-op.

DGL itself is not installed here (and the reference pins no version), so this
boundary is restated, not executed: parity unpinned at the DGL boundary.
"""
from types import SimpleNamespace

import torch


class _View:
    def __init__(self, data):
        self.data = data


class FakeGraph:
    def __init__(self, num_nodes, edges):
        """edges: {'net': (src, dst), 'cell': (src, dst)} int64 tensors."""
        self.N = int(num_nodes)
        self.ndata = {}
        self.nodes = {"pin": _View(self.ndata)}
        self.edges = {}
        self._csr = {}
        self._num_edges = {}
        for et, (src, dst) in edges.items():
            src = torch.as_tensor(src, dtype=torch.int64)
            dst = torch.as_tensor(dst, dtype=torch.int64)
            order = torch.argsort(dst, stable=True)          # in-edge order = edge-id order
            ptr = torch.zeros(self.N + 1, dtype=torch.int64)
            ptr[1:] = torch.cumsum(torch.bincount(dst, minlength=self.N), 0)
            self._csr[et] = (ptr, src[order])
            self._num_edges[et] = int(src.numel())
            self.edges[et] = _View({})

    def number_of_nodes(self):
        return self.N

    def number_of_edges(self, etype=None):
        return self._num_edges[etype]

    def to(self, device):
        return self

    def pull(self, v, message_func, reduce_func, apply_node_func=None, etype=None):
        v = torch.as_tensor(v, dtype=torch.int64)
        if v.numel() == 0:
            return
        ptr, src = self._csr[etype]
        start = ptr[v]
        deg = ptr[v + 1] - start
        field_in, msg_name = message_func.src_field, message_func.out_field
        feat = self.ndata[field_in]
        D = feat.shape[1:]
        if hasattr(reduce_func, "kind"):                      # builtin reduce
            out_name = reduce_func.out_field
            red = torch.zeros((v.numel(),) + tuple(D), dtype=feat.dtype)
            for d in torch.unique(deg).tolist():
                if d == 0:
                    continue
                sel = torch.nonzero(deg == d).squeeze(1)
                slots = start[sel][:, None] + torch.arange(d)[None, :]
                mailbox = feat[src[slots]]
                if reduce_func.kind == "mean":
                    r = mailbox.sum(1) / d
                elif reduce_func.kind == "max":
                    r = mailbox.max(1)[0]
                else:
                    r = mailbox.sum(1)
                red = red.index_copy(0, sel, r)
        else:                                                 # UDF, degree bucketing
            red, out_name = None, None
            for d in torch.unique(deg).tolist():
                if d == 0:
                    continue
                sel = torch.nonzero(deg == d).squeeze(1)
                slots = start[sel][:, None] + torch.arange(d)[None, :]
                mailbox = feat[src[slots]]
                nb = SimpleNamespace(data={k: t[v[sel]] for k, t in self.ndata.items()},
                                     mailbox={msg_name: mailbox})
                res = reduce_func(nb)
                assert len(res) == 1
                out_name, r = next(iter(res.items()))
                if red is None:
                    red = torch.zeros((v.numel(),) + tuple(r.shape[1:]), dtype=r.dtype)
                red = red.index_copy(0, sel, r)
            if red is None:                                   # every row had zero in-degree
                out_name = "h_neigh1"
                red = torch.zeros((v.numel(),) + tuple(D), dtype=feat.dtype)
        data = {k: t[v] for k, t in self.ndata.items()}
        data[out_name] = red
        res = apply_node_func(SimpleNamespace(data=data)) if apply_node_func else {out_name: red}
        for k, val in res.items():
            if k not in self.ndata:
                self.ndata[k] = torch.zeros((self.N,) + tuple(val.shape[1:]), dtype=val.dtype)
            self.ndata[k] = self.ndata[k].index_copy(0, v, val)
