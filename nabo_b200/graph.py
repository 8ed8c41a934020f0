"""``Graph``: the part of ``nabo.Graph`` (nabo/_graph.py) that sits on the hot path.

Kept verbatim: ``Graph()`` (an ``nx.Graph``), ``load_from_h5(fn, name, kind)``,
``get_mapping_score(target, ...)`` with every keyword, ``classify_target``,
``get_mapping_specificity`` / ``get_ref_specificity`` and the attributes
``refName, refNodes, refG, targetNames, targetNodes``.  The per-reference score and the
per-target cluster vote run on the GPU from the columnar graph arrays (``knn``, ``snn``)
that ``Mapping.calc_snn`` stores; the networkx structure is still populated so that
downstream networkx code keeps working (``materialize=False`` skips that for large graphs).
The rest of ``nabo.Graph`` (layouts, clustering, GML, DE helpers; nabo/_graph.py:118-554,
858-1055) is downstream analytics on the result and is out of scope.
"""
from __future__ import annotations

import os
from collections import Counter
from typing import Dict, List, Optional

import networkx as nx
import numpy as np
import pandas as pd

from . import core
from .store import Group, open_file

__all__ = ["Graph"]


class _Sample:
    """Columnar arrays of one loaded sample."""

    def __init__(self, nodes: List[str], knn: np.ndarray, snn: np.ndarray, k: int):
        self.nodes, self.knn, self.snn, self.k = nodes, knn, snn, k
        self.index = {n: i for i, n in enumerate(nodes)}


class Graph(nx.Graph):
    """Nabo's SNN graph (inherits networkx's ``Graph``)."""

    def __init__(self):
        super().__init__()
        self.refName = None
        self.refNodes: List[str] = []
        self.refG = None
        self.targetNames: List[str] = []
        self.targetNodes: Dict[str, List[str]] = {}
        self.deTestCells: List[str] = None
        self.deCtrlCells: List[str] = None
        self.clusters: Dict[str, str] = {}
        self._samples: Dict[str, _Sample] = {}
        self._refIndex: Dict[str, int] = {}

    # ------------------------------------------------------------------ loading
    def load_from_h5(self, fn: str, name: str, kind: str, materialize: Optional[bool] = None) -> None:
        """nabo/_graph.py:31-116.  Reads the columnar ``<uid>_graph`` group written by
        ``nabo_b200.Mapping`` or the reference's one-dataset-per-node layout."""
        if os.path.exists(fn) is False:
            raise IOError("ERROR: File %s does not exist" % fn)
        if kind == "reference":
            if self.refName is not None:
                raise ValueError("ERROR: A reference kind is already loaded")
        elif kind == "target":
            if name in self.targetNames:
                raise ValueError("ERROR: %s target group already present in graph" % name)
            if self.refName is None:
                raise ValueError("ERROR: Please load reference kind first")
        else:
            raise ValueError('ERROR: Kind can be either "reference" or "target"')
        try:
            h5 = open_file(fn, mode="r")
        except (IOError, OSError):
            raise IOError("ERROR: Unable to open file %s" % fn)
        if kind == "reference":
            try:
                saved_name = h5["name_stash/ref_name"][0].decode("UTF-8")
                uid = h5["name_stash/ref_name"][1].decode("UTF-8")
            except KeyError:
                raise KeyError("ERROR: Could not find stashed names in the mapping file. Make sure reference "
                               "graph has been created in the mapping file")
            if name != saved_name:
                raise KeyError("ERROR: The reference is named %s in the mapping file and not %s. Please verify "
                               "that you are trying to load right reference." % (saved_name, name))
        else:
            try:
                target_names = h5["name_stash/target_names"][:]
            except KeyError:
                raise KeyError("ERROR: Could not find stashed names in the mapping file. Make sure reference "
                               "graph has been created in the mapping file")
            uid = None
            for i in target_names:
                if i[0].decode("UTF-8") == name:
                    uid = i[1].decode("UTF-8")
            if uid is None:
                raise KeyError("ERROR: The target name not could not be found in the mapping file")
        grp = uid + "_graph"
        if grp not in h5:
            h5.close()
            raise KeyError("ERROR: Group %s not found in HDF5 file %s" % (grp, fn))
        g = h5[grp]
        ref_cells = [x.decode("UTF-8") for x in h5["ref_cells/ref_cells"][:]]
        if "knn" in g and "snn" in g:
            new_nodes, edges = self._load_columnar(g, name, kind, ref_cells)
        else:
            new_nodes, edges = self._load_per_node(g, name, kind, ref_cells)
        h5.close()
        existing = set(self.nodes()) if self.number_of_nodes() else set()
        dup = [n for n in new_nodes if n in existing]
        for n in dup:
            print("WARNING: node %s already present in the graph. Will not add." % n)
        n_edges = len(edges["u"]) if isinstance(edges, dict) else len(edges[0])
        if materialize is None:
            materialize = n_edges <= 5_000_000
        if materialize:
            import gc
            gc_was_on = gc.isenabled()
            gc.disable()            # millions of small dicts: the cyclic collector would rescan them over and over
            try:
                attrs = {"kind": kind, "name": name}
                self.add_nodes_from((n for n in new_nodes if n not in existing), **attrs)
                if isinstance(edges, dict):
                    self._bulk_add_indexed_edges(edges["names"], edges["u"], edges["v"], edges["w"], fresh=not dup)
                else:
                    a, b, w = edges
                    self.add_weighted_edges_from(zip(a, b, w))
            finally:
                if gc_was_on:
                    gc.enable()
        if kind == "reference":
            self.refName = name
            self.refNodes = [n for n in new_nodes if n not in existing]
            self._refIndex = {c + "_" + name: i for i, c in enumerate(ref_cells)}
            self.refG = self.subgraph(self.refNodes)
        else:
            self.targetNames.append(name)
            self.targetNodes[name] = [n for n in new_nodes if n not in existing]
        return None

    def _load_columnar(self, g: Group, name: str, kind: str, ref_cells: List[str]):
        nodes = [x.decode("UTF-8") for x in g["nodes"][:]]
        knn = np.asarray(g["knn"][:], dtype=np.int32)
        snn = np.asarray(g["snn"][:], dtype=np.uint8)
        k = int(g["k"][0])
        ref_suffix = g["ref_suffix"][0].decode("UTF-8") if "ref_suffix" in g else (self.refName or name)
        self._samples[name] = _Sample(nodes, knn, snn, k)
        lut = core.snn_weight_lut(k, strict=False)
        rows, cols = np.nonzero(snn > 0)
        nb = knn[rows, cols]
        w = lut[snn[rows, cols]]
        if kind == "reference":
            # undirected graph: the later add_edge wins (nx.Graph semantics of _calc_snn, :196-198):
            # targets are processed in row order, so for a pair {a, b} the weight seen from max(a, b) stays
            hi, lo = np.maximum(rows, nb), np.minimum(rows, nb)
            from_hi = rows >= nb
            order = np.lexsort((from_hi, lo, hi))           # per pair: edge seen from `lo` first, from `hi` last
            hi, lo, w = hi[order], lo[order], w[order]
            last = np.ones(len(hi), dtype=bool)
            last[:-1] = (hi[:-1] != hi[1:]) | (lo[:-1] != lo[1:])
            hi, lo, w = hi[last], lo[last], w[last]
            u, v = lo.astype(np.int64), hi.astype(np.int64)
            if "fix_edges" in g:
                fw = float(g["fix_weight"][0])
                fx = np.asarray(g["fix_edges"][:], dtype=np.int64)
                self._samples[name].fix_edges = fx
                u, v = np.concatenate([u, fx[:, 0]]), np.concatenate([v, fx[:, 1]])
                w = np.concatenate([w, np.full(len(fx), fw)])
            return nodes, {"names": nodes, "u": u, "v": v, "w": w}
        ref_names = [c + "_" + ref_suffix for c in ref_cells]
        # one name space: the sample's own nodes first, then the reference nodes
        return nodes, {"names": nodes + ref_names, "u": rows.astype(np.int64), "v": nb.astype(np.int64) + len(nodes), "w": w}

    def _load_per_node(self, g: Group, name: str, kind: str, ref_cells: List[str]):
        """Reference layout: dataset per node = rows of [neighbour name, str(weight)] (nabo/_mapping.py:265-270)."""
        nodes = [n for n in g]
        a, b, w = [], [], []
        for node in nodes:
            for j in g[node]:
                a.append(node)
                b.append(j[0].decode("UTF-8"))
                w.append(float(j[1].decode("UTF-8")))
        if kind == "target":
            ref_name = self.refName
            ridx = {c + "_" + ref_name: i for i, c in enumerate(ref_cells)}
            kmax = max(Counter(a).values()) if a else 1
            knn = np.full((len(nodes), kmax), -1, dtype=np.int32)
            wts = np.zeros((len(nodes), kmax), dtype=np.float64)
            pos = {n: i for i, n in enumerate(nodes)}
            fill = np.zeros(len(nodes), dtype=np.int64)
            for x, y, ww in zip(a, b, w):
                i = pos[x]
                knn[i, fill[i]] = ridx[y]
                wts[i, fill[i]] = ww
                fill[i] += 1
            s = _Sample(nodes, knn, None, kmax)
            s.weights = wts
            self._samples[name] = s
        return nodes, (a, b, w)

    def _bulk_add_indexed_edges(self, names, u, v, w, fresh: bool) -> None:
        """``add_weighted_edges_from`` for the edge table of a columnar graph group (u[i], v[i] index ``names``).
        With ``fresh`` (the sample's nodes were all new, so none of these edges exists yet) the adjacency dicts
        are filled per node with C-level ``dict.update`` calls - same adjacency order (edge position) and the
        same shared attribute dict per edge as networkx builds - instead of one interpreter round trip per edge
        (3 M edges: 3.5 s -> ~1 s).  Otherwise it goes through networkx itself."""
        names_arr = np.empty(len(names), dtype=object)
        names_arr[:] = names
        wl = np.asarray(w, dtype=np.float64).tolist()
        if not fresh:
            self.add_weighted_edges_from(zip(names_arr[u].tolist(), names_arr[v].tolist(), wl))
            return
        n_e = len(wl)
        dds = np.empty(n_e, dtype=object)
        dds[:] = [{"weight": x} for x in wl]
        node = np.concatenate([u, v])
        other = np.concatenate([v, u])
        pos = np.concatenate([2 * np.arange(n_e), 2 * np.arange(n_e) + 1])       # networkx writes adj[u][v], then adj[v][u]
        order = np.lexsort((pos, node))
        node, other_names, dd = node[order], names_arr[other[order]], dds[np.concatenate([np.arange(n_e)] * 2)[order]]
        bounds = np.flatnonzero(np.r_[True, node[1:] != node[:-1], True]) if len(node) else np.zeros(1, dtype=np.int64)
        adj = self._adj
        for lo, hi in zip(bounds[:-1].tolist(), bounds[1:].tolist()):
            adj[names[node[lo]]].update(zip(other_names[lo:hi].tolist(), dd[lo:hi].tolist()))

    # ------------------------------------------------------------------ mapping score
    def get_mapping_score(self, target: str, min_weight: float = 0, min_score: float = 0, weighted: bool = True,
                          by_cluster: bool = False, sorted_names_only: bool = False, top_n_only: int = None,
                          all_nodes: bool = True, score_multiplier: int = 1000, ignore_nodes: List[str] = None,
                          include_nodes: List[str] = None, remove_suffix: bool = False, verbose: bool = False):
        """nabo/_graph.py:555-697; the per-reference accumulation (:643-653) runs on the GPU."""
        if by_cluster:
            if not self.clusters:
                raise ValueError('ERROR: Calculate clusters first using "make_clusters" or import clusters using '
                                 '"import_clusters"')
        if target not in self.targetNames:
            raise ValueError("ERROR: %s not present in graph" % target)
        if ignore_nodes is not None and include_nodes is not None:
            raise ValueError("ERROR: PLease provide only one of either 'ignore_nodes' or 'include_nodes' at a time")
        s = self._samples[target]
        tset = s.index
        ignore = [] if ignore_nodes is None else [n for n in ignore_nodes if n in tset]
        include = list(self.targetNodes[target]) if include_nodes is None else [n for n in include_nodes if n in tset]
        include = list(set(include).difference(ignore))
        if len(include) == 0:
            raise ZeroDivisionError("division by zero")          # len(include_nodes) == 0 upstream (:652)
        mask = np.zeros(len(s.nodes), dtype=np.uint8)
        mask[[tset[n] for n in include]] = 1
        m = len(self._refIndex)
        if s.snn is not None:
            vals = core.mapping_scores(s.knn, s.snn, m, s.k, include=mask, min_weight=float(min_weight),
                                       min_score=-np.inf, weighted=bool(weighted),
                                       score_multiplier=float(score_multiplier))
        else:                                                    # graph loaded from the per-node layout
            vals = _scores_from_weights(s.knn, s.weights, m, mask, min_weight, weighted, score_multiplier)
        if verbose:
            deg = np.zeros(m, dtype=np.int64)
            valid = (s.knn >= 0) & ((s.snn > 0) if s.snn is not None else (s.knn >= 0)) & (mask[:, None] > 0)
            np.add.at(deg, s.knn[valid], 1)
            print("INFO: The bipartite graph has %d edges" % int(valid.sum()))
            print("INFO: Mapping calculated against %d %s nodes" % (len(include), target))
            print("INFO: %d reference nodes do not connect with any target node" % int((deg == 0).sum()))
            print("INFO: %d target nodes do not connect with any reference node"
                  % int((~valid.any(axis=1) & (mask > 0)).sum()))
        score = {n: float(vals[self._refIndex[n]]) for n in self.refNodes}

        if by_cluster:
            cluster_dict = self.clusters
            cluster_values = {x: [] for x in set(cluster_dict.values())}
            na_cluster_score = []
            for node in score:
                try:
                    cluster_values[cluster_dict[node]].append(score[node])
                except KeyError:
                    na_cluster_score.append(score[node])
            if len(na_cluster_score) > 0:
                if "NA" not in cluster_values:
                    cluster_values["NA"] = []
                else:
                    print("WARNING: 'NA' cluster already exists. Appending value to it")
                cluster_values["NA"].extend(na_cluster_score)
            return cluster_values
        if sorted_names_only:
            if top_n_only is not None:
                if top_n_only > len(score):
                    raise ValueError("ERROR: Value of top_n_only should be less than total number of nodes in "
                                     "reference graph")
                retval = [x[0] for x in sorted(score.items(), key=lambda x: x[1])][::-1][:top_n_only]
            else:
                ms = {k: v for k, v in score.items() if v >= min_score}
                retval = [x[0] for x in sorted(ms.items(), key=lambda x: x[1])][::-1]
            return [x.rsplit("_", 1)[0] for x in retval] if remove_suffix else retval
        if not all_nodes:
            retval = {k: v for k, v in score.items() if v >= min_score}
        else:
            retval = {k: v if v >= min_score else 0 for k, v in score.items()}
        return [x.rsplit("_", 1)[0] for x in retval] if remove_suffix else retval

    # ------------------------------------------------------------------ classification
    def import_clusters(self, cluster_dict: Dict[str, str]) -> None:
        """Attach reference cluster labels (node name -> label), cf. nabo/_graph.py:398-432."""
        self.clusters = {k: v for k, v in cluster_dict.items() if k in self._refIndex}

    def classify_target(self, target: str, weight_frac: float = 0.5, min_degree: int = 2, min_weight: float = 0.1,
                        cluster_dict: Dict[str, int] = None, na_label: str = "NA", ret_counts: bool = False) -> dict:
        """nabo/_graph.py:722-792 on the GPU (per-target cluster vote)."""
        if cluster_dict is None:
            if not self.clusters:
                raise ValueError("ERROR: Please make sure that clusters are set for each reference node")
            cluster_dict = self.clusters
        s = self._samples[target]
        if s.snn is None:
            raise ValueError("ERROR: classify_target needs a graph stored in the columnar layout")
        labels = sorted(set(cluster_dict.values()), key=str)
        lid = {l: i for i, l in enumerate(labels)}
        ref_labels = np.full(len(self._refIndex), -1, dtype=np.int32)
        for node, lab in cluster_dict.items():
            if node in self._refIndex:
                ref_labels[self._refIndex[node]] = lid[lab]
        out = core.classify_targets(s.knn, s.snn, ref_labels, len(labels), s.k, weight_frac=weight_frac,
                                    min_degree=min_degree, min_weight=min_weight)
        classified = [labels[i] if i >= 0 else na_label for i in out]
        if ret_counts:
            counts = Counter(classified)
            if na_label not in counts:
                counts[na_label] = 0
            for i in set(cluster_dict.values()):
                if i not in counts:
                    counts[i] = 0
            return counts
        return dict(zip(s.nodes, classified))


    # ------------------------------------------------------------------ mapping specificity
    def _ref_csr(self):
        """Symmetric CSR adjacency of the reference graph over reference-cell indices."""
        import scipy.sparse as sp
        m = len(self._refIndex)
        s = self._samples.get(self.refName)
        if s is not None and s.snn is not None:
            rows, cols = np.nonzero(s.snn > 0)
            a, b = rows.astype(np.int64), s.knn[rows, cols].astype(np.int64)
            fx = getattr(s, "fix_edges", None)
            if fx is not None and len(fx):
                a, b = np.concatenate([a, fx[:, 0]]), np.concatenate([b, fx[:, 1]])
        else:                                              # per-node layout: walk the networkx graph
            ea = [(self._refIndex[x], self._refIndex[y]) for x, y in self.refG.edges()]
            a = np.array([e[0] for e in ea], dtype=np.int64)
            b = np.array([e[1] for e in ea], dtype=np.int64)
        keep = a != b
        a, b = a[keep], b[keep]
        adj = sp.coo_matrix((np.ones(2 * len(a), np.int8), (np.concatenate([a, b]), np.concatenate([b, a]))),
                            shape=(m, m)).tocsr()
        adj.sum_duplicates()
        return adj.indptr.astype(np.int64), adj.indices.astype(np.int32)

    def get_mapping_specificity(self, target_name: str, fill_na: bool = True) -> Dict[str, float]:
        """nabo/_graph.py:794-824 on the GPU: mean unweighted shortest-path length in the reference graph
        between all pairs of reference nodes a target node maps to (one bit-parallel multi-source BFS per
        target instead of one networkx BFS per pair).  NaN for targets with fewer than two mapped nodes
        (replaced by the largest value when ``fill_na``); ``networkx.NetworkXNoPath`` when two mapped nodes
        are not connected, as upstream."""
        s = self._samples[target_name]
        if s.snn is not None:
            cnt = (s.snn > 0).astype(np.uint8)
        else:
            cnt = (s.knn >= 0).astype(np.uint8)
        indptr, indices = self._ref_csr()
        mean, connected = core.mapping_specificity(indptr, indices, s.knn, cnt)
        if not bool(np.all(connected)):
            bad = s.nodes[int(np.nonzero(~connected)[0][0])]
            raise nx.NetworkXNoPath("No path between two of the reference nodes that %s maps to." % bad)
        path_lengths = {n: float(v) for n, v in zip(s.nodes, mean)}
        if fill_na:
            max_val = max(path_lengths.values())            # NaN-propagating exactly like the builtin upstream
            return pd.Series(path_lengths).fillna(max_val).to_dict()
        return path_lengths

    def get_ref_specificity(self, target: str, target_values: Dict[str, float],
                            incl_unmapped: bool = False) -> Dict[str, float]:
        """nabo/_graph.py:826-857: mean specificity of the target nodes mapped to each reference node
        (adjacency order = target order, as upstream)."""
        s = self._samples[target]
        mask = (s.snn > 0) if s.snn is not None else (s.knn >= 0)
        rows, cols = np.nonzero(mask)
        refs = s.knn[rows, cols]
        order = np.argsort(refs, kind="stable")             # per reference node: its targets in target order
        refs, rows = refs[order], rows[order]
        vals = np.array([target_values[n] for n in s.nodes], dtype=np.float64)[rows]
        bounds = np.flatnonzero(np.r_[True, refs[1:] != refs[:-1], True])
        names = {i: n for n, i in self._refIndex.items()}
        out: Dict[str, float] = {}
        have = set()
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            r = int(refs[lo])
            have.add(r)
            out[names[r]] = np.mean(vals[lo:hi]) if hi - lo > 1 else vals[lo]
        if incl_unmapped:
            for n, i in self._refIndex.items():
                if i not in have:
                    out[n] = 0
        return {n: out[n] for n in self.refNodes if n in out}


def _scores_from_weights(knn, weights, n_ref, mask, min_weight, weighted, mult):
    """Per-node-layout graphs carry weights, not SNN counts: map weight -> rank table -> GPU kernel."""
    exists = knn >= 0                                   # an edge is an entry of the node's dataset, whatever its weight
    vals = np.unique(weights[exists])
    lut = np.concatenate([[0.0], vals])
    if len(lut) > 255:
        raise ValueError("ERROR: more than 254 distinct edge weights; store the graph in the columnar layout")
    cnt = (np.searchsorted(vals, weights) + 1).astype(np.uint8)
    cnt[~exists] = 0
    import torch
    from . import _lib
    import ctypes as C
    core.require_device()
    td, cd = core._dev(knn, torch.int32), core._dev(cnt, torch.uint8)
    ld, md = core._dev(lut, torch.float64), core._dev(mask, torch.uint8)
    n, kk = td.shape
    out = torch.empty(n_ref, dtype=torch.float64, device=td.device)
    L = _lib.lib()
    ws = torch.empty(int(L.nabo_scores_workspace_bytes(n, kk, n_ref)), dtype=torch.uint8, device=td.device)
    _lib.check(L.nabo_mapping_scores(core._ptr(td), core._ptr(cd), core._ptr(ld), n, kk, int(n_ref), core._ptr(md),
                                     int(mask.sum()), float(min_weight), 1 if weighted else 0, float(mult),
                                     float("-inf"), core._ptr(out), core._ptr(ws), ws.numel(),
                                     C.c_void_p(core._stream())), "mapping_scores")
    return out.cpu().numpy()
