"""Reading the reference's processed split files without torch-geometric.

`EUPPBench.process` (utils/dataset.py:174-182) writes `torch.save((data, slices), path)` where `data` is a pickled
`torch_geometric.data.Data` - a plain object whose `__dict__` holds `_store`, a `torch_geometric.data.storage.GlobalStorage`
with the attribute dict in `_mapping` (PyG 2.x) - and `slices` a dict of offset tensors.  torch-geometric is not installable
here, so the classes named in the pickle are replaced by attribute bags while loading; nothing of PyG is executed.
"""
from __future__ import annotations

import pickle


class _Bag:
    """Stand-in for any torch_geometric class found in a pickle: keeps the pickled state as attributes."""

    def __init__(self, *args, **kwargs):
        self._args, self._kwargs = args, kwargs

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__["_state"] = state

    def mapping(self) -> dict:
        """The attribute dict of a PyG Data / storage object, whatever the nesting."""
        d = self.__dict__
        if "_mapping" in d:
            return d["_mapping"]
        if "_store" in d and isinstance(d["_store"], _Bag):
            return d["_store"].mapping()
        return {k: v for k, v in d.items() if not k.startswith("_")}


_stubs = {}


class Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "torch_geometric" or module.startswith("torch_geometric."):
            key = (module, name)
            if key not in _stubs:
                _stubs[key] = type(name, (_Bag,), {"__module__": "raincast_gnn_b200.pyg_compat.unpickle"})
            return _stubs[key]
        return super().find_class(module, name)


def load(file, **kw):
    return Unpickler(file, **kw).load()


def attribute_dict(data) -> dict:
    """dict of attributes of a loaded `data` object: our own dict format, a stand-in bag, or anything with attributes."""
    if isinstance(data, dict):
        return data
    if isinstance(data, _Bag):
        return data.mapping()
    return {k: getattr(data, k) for k in ("x", "ensemble", "y", "edge_index", "edge_attr") if hasattr(data, k)}
