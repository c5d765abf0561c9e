"""Stand-in for python-igraph, written for this repository (test infrastructure; python-igraph is not installed in the
build container or on the GPU box). It provides only the calls the reference's mapping stage makes
(bin/ntlink_pair.py:140-155,267-305,498-506; bin/ntlink_utils.py:36-62): Graph(directed), add_vertices, add_edges, get_eid,
vs()/vs.find, es() with attribute columns, copy, delete_edges, InternalError. All look-ups are O(1) so that the reference
arm of bench.py is not slowed down by the stand-in. `make -C oracle ref` installs it as oracle/_ref/igraph.py next to the
unmodified reference files."""


class InternalError(Exception):
    pass


class _Vertex:
    __slots__ = ("g", "index")

    def __init__(self, g, i):
        self.g, self.index = g, i

    def __getitem__(self, key):
        if key != "name":
            raise KeyError(key)
        return self.g._names[self.index]


class _VertexSeq:
    def __init__(self, g):
        self.g = g

    def __call__(self):
        return self

    def __len__(self):
        return len(self.g._names)

    def __getitem__(self, i):
        return _Vertex(self.g, range(len(self.g._names))[i])

    def __iter__(self):
        return (_Vertex(self.g, i) for i in range(len(self.g._names)))

    def find(self, name=None, **kw):
        name = kw.get("name", name)
        try:
            return _Vertex(self.g, self.g._idx[name])
        except KeyError:
            raise ValueError(f"no such vertex: {name}") from None


class _Edge:
    __slots__ = ("g", "index")

    def __init__(self, g, i):
        self.g, self.index = g, i

    @property
    def source(self):
        return self.g._edges[self.index][0]

    @property
    def target(self):
        return self.g._edges[self.index][1]

    def __getitem__(self, key):
        return self.g._attrs[key][self.index]


class _EdgeSeq:
    def __init__(self, g):
        self.g = g

    def __len__(self):
        return len(self.g._edges)

    def __iter__(self):
        return (_Edge(self.g, i) for i in range(len(self.g._edges)))

    def __getitem__(self, key):
        if isinstance(key, str):
            return self.g._attrs[key]
        return _Edge(self.g, range(len(self.g._edges))[key])

    def __setitem__(self, key, values):
        values = list(values)
        if len(values) != len(self.g._edges):
            raise ValueError("attribute column length != number of edges")
        self.g._attrs[key] = values


class Graph:
    def __init__(self, directed=True):
        self._names, self._idx, self._edges, self._eid, self._attrs = [], {}, [], {}, {}
        self.vs = _VertexSeq(self)

    def add_vertices(self, names):
        for n in names:
            self._idx[n] = len(self._names)
            self._names.append(n)

    def add_edges(self, pairs):
        for s, t in pairs:
            e = (self._idx[s] if not isinstance(s, int) else s, self._idx[t] if not isinstance(t, int) else t)
            self._eid.setdefault(e, len(self._edges))
            self._edges.append(e)

    def get_eid(self, s, t):
        try:
            return self._eid[(self._idx[s] if not isinstance(s, int) else s, self._idx[t] if not isinstance(t, int) else t)]
        except KeyError:
            raise InternalError("no such edge") from None

    def es(self):
        return _EdgeSeq(self)

    def vcount(self):
        return len(self._names)

    def ecount(self):
        return len(self._edges)

    def copy(self):
        g = Graph()
        g._names, g._idx = list(self._names), dict(self._idx)
        g._edges, g._eid = list(self._edges), dict(self._eid)
        g._attrs = {k: list(v) for k, v in self._attrs.items()}
        return g

    def delete_edges(self, idxs):
        drop = set(idxs)
        keep = [i for i in range(len(self._edges)) if i not in drop]
        self._edges = [self._edges[i] for i in keep]
        self._attrs = {k: [v[i] for i in keep] for k, v in self._attrs.items()}
        self._eid = {}
        for i, e in enumerate(self._edges):
            self._eid.setdefault(e, i)
