"""Minimal HeteroData: attribute stores per node/edge type (used at graph.py:78-87, mapper.py:141-145 of the reference)."""


class _Store(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    @property
    def num_nodes(self):
        if "num_nodes" in self:
            return self["num_nodes"]
        for v in self.values():
            if hasattr(v, "shape"):
                return v.shape[0]
        return None

    def node_attrs(self):
        return [k for k in self.keys()]

    def edge_attrs(self):
        return [k for k in self.keys()]


class HeteroData:
    def __init__(self):
        object.__setattr__(self, "_stores", {})

    def __getitem__(self, key):
        st = self._stores
        if key not in st:
            st[key] = _Store()
        return st[key]

    def __setitem__(self, key, value):
        self._stores[key] = value

    def __bool__(self):
        return True

    @property
    def node_types(self):
        return [k for k in self._stores if isinstance(k, str)]

    @property
    def edge_types(self):
        return [k for k in self._stores if isinstance(k, tuple)]

    def node_items(self):
        return [(k, v) for k, v in self._stores.items() if isinstance(k, str)]

    def edge_items(self):
        return [(k, v) for k, v in self._stores.items() if isinstance(k, tuple)]

    def to(self, *a, **k):
        for s in self._stores.values():
            for key, v in list(s.items()):
                if hasattr(v, "to"):
                    s[key] = v.to(*a, **k)
        return self
