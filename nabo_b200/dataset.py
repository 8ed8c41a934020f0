"""``Dataset``: the slice of ``nabo.Dataset`` (nabo/_dataset.py) that feeds the hot path.

On the path (GPU): ``get_scaled_values`` (:846-915) and ``transform_pca`` (:985-1033) -
scaling and the projection of target counts into the reference's HVG/PCA space.
Inputs to the path (host, same arithmetic as upstream so that mu/sigma are bit-identical):
``set_sf`` (:548-592), ``set_gene_stats`` (:594-637), ``get_scaling_params`` (:814-844),
``fit_ipca`` (:917-983, scikit-learn IncrementalPCA with the reference's batch schedule).
QC plots, HVG selection with lowess, marker tests etc. are out of scope (SURVEY.md section 2,
rows 10-14): pass the gene list you want to ``fit_ipca``.

File layout: the reference's (``names/{genes,cells}``, ``cell_data/<cell>`` structured
(idx, val) arrays, ``processed_data/*``; nabo/_io.py:103-115) or the columnar
``cell_csr/{indptr,idx,val}`` written by ``write_dataset`` below.
"""
from __future__ import annotations

from typing import Dict, Generator, List, Optional, Tuple

import numpy as np
import pandas as pd

from . import core
from .store import open_file

__all__ = ["Dataset", "write_dataset"]


def write_dataset(fn: str, counts: np.ndarray, cells: List[str], genes: List[str]) -> None:
    """Create a dataset file from a dense (cells x genes) count matrix (test / bench helper;
    MTX/CSV ingest, nabo/_io.py, is out of scope)."""
    import scipy.sparse as sp
    csr = sp.csr_matrix(np.asarray(counts))
    h5 = open_file(fn, mode="w")
    ng = h5.create_group("names")
    ng.create_dataset("genes", data=np.array([g.upper().encode("ascii") for g in genes]))
    ng.create_dataset("cells", data=np.array([c.upper().encode("ascii") for c in cells]))
    g = h5.create_group("cell_csr")
    g.create_dataset("indptr", data=csr.indptr.astype(np.int64))
    g.create_dataset("idx", data=csr.indices.astype(np.int32))
    g.create_dataset("val", data=csr.data.astype(np.float32))
    h5.close()


class Dataset:
    """:param h5_fn: dataset file; :param force_recalc: drop cached ``processed_data``.
    ``mito_patterns`` / ``ribo_patterns`` are accepted for signature compatibility."""

    def __init__(self, h5_fn: str, mito_patterns: List[str] = None, ribo_patterns: List[str] = None,
                 force_recalc: bool = False):
        self.h5Fn = h5_fn
        self._mito_patterns = mito_patterns if mito_patterns is not None else ["^MT-"]
        self._ribo_patterns = ribo_patterns if ribo_patterns is not None else ["^RPS", "^RPL", "^MRPS", "^MRPL"]
        self._recalc = force_recalc
        self.cells: List[str] = None
        self.genes: List[str] = None
        self.rawNCells: int = None
        self.rawNGenes: int = None
        self.cellIdx: Dict[str, int] = None
        self.geneIdx: Dict[str, int] = None
        self.keepCellsIdx = None
        self.keepGenesIdx = None
        self.sf: np.ndarray = None
        self.geneStats: pd.DataFrame = None
        self.hvgList = None
        self.ipca = None
        self._csr = None
        self._load_info()

    # ------------------------------------------------------------------ loading
    def _load_info(self):
        h5 = open_file(self.h5Fn, mode="a")
        try:
            self.cells = [x.decode("UTF-8") for x in h5["names"]["cells"][:]]    # never sorted (_dataset.py:93)
            self.genes = [x.decode("UTF-8") for x in h5["names"]["genes"][:]]
        except KeyError:
            self.cells, self.genes = [], []
            h5.close()
            raise IOError("FATAL ERROR: Could not extract gene/cell names from the H5 file. Please make sure "
                          "that file was generated using Nabo's IO functions ")
        self.rawNCells, self.rawNGenes = len(self.cells), len(self.genes)
        self.cellIdx = {x: n for n, x in enumerate(self.cells)}
        self.geneIdx = {x: n for n, x in enumerate(self.genes)}
        if self._recalc is True and "processed_data" in h5:
            print("INFO: Deleting existing `processed_data` group from HDF5 file", flush=True)
            del h5["processed_data"]
        pd_grp = h5["processed_data"] if "processed_data" in h5 else h5.create_group("processed_data")
        if "keep_cells_idx" in pd_grp:
            self.keepCellsIdx = np.array(list(pd_grp["keep_cells_idx"][:]))
            print("INFO: Cached filtered cells loaded", flush=True)
        else:
            self.keepCellsIdx = np.array(list(range(self.rawNCells)))
        if "keep_genes_idx" in pd_grp:
            self.keepGenesIdx = np.array(list(pd_grp["keep_genes_idx"][:]))
            print("INFO: Cached filtered genes loaded", flush=True)
        else:
            self.keepGenesIdx = np.array(list(range(self.rawNGenes)))
        if "sf" in pd_grp:
            self.sf = np.asarray(pd_grp["sf"][:])
            print("INFO: Cached cell size factors loaded", flush=True)
        else:
            self.sf = np.ones(self.rawNCells, dtype=np.float32)
        if "hvg_list" in pd_grp:
            self.hvgList = [x.decode("UTF-8") for x in pd_grp["hvg_list"]]
            print("INFO: Loaded cached HVG names", flush=True)
        h5.close()

    def _counts_csr(self):
        """(indptr int64, idx int32, val float32) over all cells / all genes, in self.cells order."""
        if self._csr is None:
            h5 = open_file(self.h5Fn, mode="r")
            if "cell_csr" in h5:
                g = h5["cell_csr"]
                self._csr = (np.asarray(g["indptr"][:], np.int64), np.asarray(g["idx"][:], np.int32),
                             np.asarray(g["val"][:], np.float32))
            else:
                ip, ix, vl = [0], [], []
                cd = h5["cell_data"]
                for c in self.cells:
                    d = cd[c]
                    ix.append(np.asarray(d["idx"], dtype=np.int32))
                    vl.append(np.asarray(d["val"], dtype=np.float32))      # a[d['idx']] = d['val'] on a float32 array
                    ip.append(ip[-1] + len(ix[-1]))
                self._csr = (np.array(ip, np.int64), np.concatenate(ix) if ix else np.zeros(0, np.int32),
                             np.concatenate(vl) if vl else np.zeros(0, np.float32))
            h5.close()
        return self._csr

    def _dense(self, cell_rows: np.ndarray, gene_cols: Optional[np.ndarray] = None) -> np.ndarray:
        import scipy.sparse as sp
        ip, ix, vl = self._counts_csr()
        m = sp.csr_matrix((vl, ix, ip), shape=(self.rawNCells, self.rawNGenes))[cell_rows]
        if gene_cols is not None:
            m = m[:, gene_cols]
        return np.asarray(m.todense(), dtype=np.float32)

    # ------------------------------------------------------------------ inputs of the path (host)
    def set_sf(self, sf: Dict[str, float] = None, size_scale: float = 1000.0, all_genes: bool = False) -> None:
        """nabo/_dataset.py:548-592: sf_i = size_scale / sum of the cell's (kept-gene) counts, float32."""
        try:
            size_scale = float(size_scale)
        except TypeError:
            raise TypeError("ERROR: size_scale parameter should have a float value. E.x. not 1 but 1.0")
        if sf is not None:
            for i in sf:
                self.sf[self.cellIdx[i]] = size_scale / sf[i]
        else:
            ip, ix, vl = self._counts_csr()
            # float32 pairwise sum of the dense (kept-)gene vector of every cell, as `temp[...].sum()` upstream:
            # one GPU thread per cell walks NumPy's summation tree (nabo_sparse_row_stats)
            if all_genes:
                pos = np.arange(self.rawNGenes, dtype=np.int32)
                n_dense = self.rawNGenes
            else:
                pos = np.full(self.rawNGenes, -1, dtype=np.int32)
                pos[self.keepGenesIdx] = np.arange(len(self.keepGenesIdx), dtype=np.int32)
                n_dense = len(self.keepGenesIdx)
            tot = core.sparse_row_stats(ip, ix, vl, pos, n_dense, moments=False)["sum"]     # float32, as temp.sum()
            tot[tot == 0] = 1
            self.sf = (size_scale / tot).astype(np.float32)         # float32 division, like `size_scale / temp_sf`
        h5 = open_file(self.h5Fn, mode="a")
        grp = h5["processed_data"]
        if "sf" in grp:
            del grp["sf"]
        grp.create_dataset("sf", data=self.sf)
        h5.close()
        return None

    def set_gene_stats(self) -> None:
        """nabo/_dataset.py:594-637: per kept gene, float32 mean / non-zero mean / population variance of
        the size-factor-normalised values over kept cells, reduced on the GPU in NumPy's float32 pairwise
        order (nabo_sparse_row_stats), so the statistics - and hence mu, sigma - are bit-identical to upstream."""
        import scipy.sparse as sp
        ip, ix, vl = self._counts_csr()
        csc = sp.csr_matrix((vl, ix, ip), shape=(self.rawNCells, self.rawNGenes)).tocsc()
        csc.sort_indices()
        pos = np.full(self.rawNCells, -1, dtype=np.int32)
        pos[self.keepCellsIdx] = np.arange(len(self.keepCellsIdx), dtype=np.int32)
        st = core.sparse_row_stats(csc.indptr.astype(np.int64), csc.indices.astype(np.int32),
                                   csc.data.astype(np.float32), pos, len(self.keepCellsIdx),
                                   scale=self.sf[self.keepCellsIdx])
        keep_genes = set(int(x) for x in self.keepGenesIdx)
        stats = {}
        for i in range(self.rawNGenes):
            gene = self.genes[i]
            if i in keep_genes and st["npos"][i] > 0:
                stats[gene] = {"m": st["mean"][i], "nzm": st["nzmean"][i], "variance": st["var"][i],
                               "valid_gene": True, "ncells": st["npos"][i]}
            else:
                stats[gene] = {"valid_gene": False}
        df = pd.DataFrame(stats).T
        for c in ("m", "nzm", "variance"):
            df[c] = pd.to_numeric(df[c]).astype(np.float64)
            df[c] = df[c].fillna(df[c].min())
        df["ncells"] = pd.to_numeric(df["ncells"]).fillna(0).astype(np.float64)
        df["valid_gene"] = df["valid_gene"].astype(bool)
        self.geneStats = df
        return None

    def get_scaling_params(self, genes: List[str] = None, only_valid: bool = True) -> pd.DataFrame:
        """nabo/_dataset.py:814-844."""
        if self.geneStats is None:
            self.set_gene_stats()
        params = pd.DataFrame({"mu": self.geneStats.m.values, "sigma": np.sqrt(self.geneStats.variance.values),
                               "genes": self.geneStats.index}).set_index("genes")
        if only_valid:
            valid_genes = {x: None for x in self.geneStats[self.geneStats.valid_gene].index}
        else:
            valid_genes = {x: None for x in self.geneStats.index}
        goi_list = [x for x in valid_genes] if genes is None else [x for x in genes if x in valid_genes]
        if len(goi_list) == 0:
            raise ValueError("None of the input genes are valid! Genes should be valid as given in geneStats "
                             "attribute")
        return params.reindex(goi_list)

    def _gene_order(self, scaling_params: pd.DataFrame, fill_missing: bool) -> np.ndarray:
        goi = np.empty(len(scaling_params.index), dtype=np.int32)
        missing = 0
        for n, x in enumerate(scaling_params.index):
            if x not in self.geneIdx:
                if fill_missing is False:
                    raise KeyError("ERROR: Gene name %s not found! It may that there are other gene names in "
                                   "'scaling_params' that are also not present in this Dataset. One can try to "
                                   "intersect gene names, or set 'fill_missing' to True (not recommended)." % x)
                goi[n] = -1
                missing += 1
            else:
                goi[n] = self.geneIdx[x]
        if missing > 0:
            print("WARNING: %d out %d genes are missing in this dataset" % (missing, len(goi)))
        return goi

    # ------------------------------------------------------------------ the path (GPU)
    def get_scaled_values(self, scaling_params: pd.DataFrame, tqdm_msg: str = "", disable_tqdm: bool = False,
                          fill_missing: bool = False, chunk: int = 4096) -> \
            Generator[Tuple[str, np.ndarray], None, bool]:
        """nabo/_dataset.py:846-915: yields (cell, ((a * sf) - mu) / sigma) in the gene order of
        ``scaling_params``; computed on the GPU a chunk of cells at a time."""
        mu = scaling_params["mu"].values.astype(np.float64)
        sigma = scaling_params["sigma"].values.astype(np.float64)
        goi = self._gene_order(scaling_params, fill_missing)
        keep = np.asarray(self.keepCellsIdx)
        for s in range(0, len(keep), chunk):
            rows = keep[s:s + chunk]
            z = core.scale_counts(self._dense(rows), goi, self.sf[rows].astype(np.float32), mu, sigma)
            for r, zi in zip(rows, z):
                yield self.cells[r], zi
        return True

    def fit_ipca(self, genes: List[str], n_comps: int = 100, batch_size: int = None,
                 disable_tqdm: bool = False, *, method: str = "sklearn") -> None:
        """nabo/_dataset.py:917-983.  ``method='sklearn'`` (default): scikit-learn IncrementalPCA over the
        reference's equal-size batches on the host - the upstream fit, bit-compatible.  ``method='device'``: the
        exact PCA of the same scaled values, moments and eigen-decomposition on the GPU (``nabo_b200.pca``); equal
        to the incremental fit at the level of the subspace."""
        if method not in ("sklearn", "device"):
            raise ValueError("ERROR: method must be 'sklearn' or 'device'")
        if method == "device":
            return self._fit_pca_device(genes, n_comps)
        from sklearn.decomposition import IncrementalPCA

        def make_eq_bins(n, bs):
            a = n // bs
            b = n // a
            c = n % a
            for i in range(a):
                yield b + 1 if i < c else b

        n_comps = int(n_comps)
        if len(genes) < n_comps:
            n_comps = len(genes)
            print("WARNING: Number of components were reset to number of features i.e. %d" % n_comps)
        if n_comps > len(self.keepCellsIdx):
            n_comps = len(self.keepCellsIdx) - 1
            print("WARNING: Number of components were reset to number of cells - 1 i.e. %d" % n_comps)
        if batch_size is None or batch_size < n_comps:
            batch_size = n_comps * 2
        if batch_size > len(self.keepCellsIdx):
            batch_size = len(self.keepCellsIdx)
        scaling_params = self.get_scaling_params(genes)
        self.ipca = IncrementalPCA(n_components=n_comps)
        cache = []
        sizer = make_eq_bins(len(self.keepCellsIdx), batch_size)
        cur = next(sizer)
        for _, a in self.get_scaled_values(scaling_params, disable_tqdm=disable_tqdm):
            cache.append(a)
            if len(cache) == cur:
                self.ipca.partial_fit(np.array(cache))
                cache = []
                try:
                    cur = next(sizer)
                except StopIteration:
                    pass
        if len(cache) > 0:
            print("WARNING: Not all cells were processed! This is a bug. Please report it to the authors.")
        self.ipca.genes = list(scaling_params.index)
        return None

    def _fit_pca_device(self, genes: List[str], n_comps: int, chunk: int = 8192) -> None:
        import torch
        from .pca import DevicePCA, MomentAccumulator
        n_comps = int(n_comps)
        if len(genes) < n_comps:
            n_comps = len(genes)
            print("WARNING: Number of components were reset to number of features i.e. %d" % n_comps)
        if n_comps > len(self.keepCellsIdx):
            n_comps = len(self.keepCellsIdx) - 1
            print("WARNING: Number of components were reset to number of cells - 1 i.e. %d" % n_comps)
        scaling_params = self.get_scaling_params(genes)
        mu = torch.from_numpy(scaling_params["mu"].values.astype(np.float64)).cuda()
        sigma = torch.from_numpy(scaling_params["sigma"].values.astype(np.float64)).cuda()
        goi = torch.from_numpy(self._gene_order(scaling_params, False)).cuda()
        keep = np.asarray(self.keepCellsIdx)
        acc = MomentAccumulator(len(scaling_params), mu.device)
        for s in range(0, len(keep), chunk):
            rows = keep[s:s + chunk]
            z = core.scale_counts(torch.from_numpy(self._dense(rows)).cuda(), goi,
                                  torch.from_numpy(self.sf[rows].astype(np.float32)).cuda(), mu, sigma)
            acc.add(z)
        self.ipca = DevicePCA(n_comps).fit_moments(acc)
        self.ipca.genes = list(scaling_params.index)
        return None

    def transform_pca(self, out_file: str, pca_group_name: str, transformer, scaling_params: pd.DataFrame,
                      disable_tqdm: bool = False, fill_missing: bool = False, chunk: int = 65536) -> None:
        """nabo/_dataset.py:985-1033: project every kept cell with the reference's scaling parameters and
        PCA model (``transformer.components_``, ``.mean_``) and store ``<group>/<cell>`` float64 vectors
        (one columnar RowGroup).  Scaling + projection are one fused CUDA kernel per chunk."""
        if transformer is None:
            raise ValueError("ERROR: None value found for transformer. Please make sure that the PCA was fitted")
        if scaling_params is None:
            raise ValueError("ERROR: scaling_params need to be a DataFrame")
        if getattr(transformer, "whiten", False):
            raise ValueError("ERROR: whitened PCA models are not supported (the reference never whitens)")
        try:
            h5 = open_file(out_file, mode="a")
        except Exception:
            raise IOError("ERROR: Could not open file %s" % out_file)
        if pca_group_name in h5:
            del h5[pca_group_name]
        try:
            goi = self._gene_order(scaling_params, fill_missing)
        except KeyError as ke:
            h5.close()
            raise KeyError(ke)
        mu = scaling_params["mu"].values.astype(np.float64)
        sigma = scaling_params["sigma"].values.astype(np.float64)
        comps = np.ascontiguousarray(transformer.components_, dtype=np.float64)
        mean = np.ascontiguousarray(transformer.mean_, dtype=np.float64)
        keep = np.asarray(self.keepCellsIdx)
        ip, ix, vl = self._counts_csr()
        pos = np.full(self.rawNGenes, -1, dtype=np.int32)
        present = goi[goi >= 0]
        pos[present] = np.nonzero(goi >= 0)[0].astype(np.int32)
        out = np.empty((len(keep), comps.shape[0]), dtype=np.float64)
        # a gene listed twice in scaling_params has ONE column in the dataset but two model positions (upstream's
        # a[goi] copies it to both, _dataset.py:905-911); the CSR kernel's gene -> position map holds one position
        # per gene, so such a model goes through the dense gather, which takes any goi
        duplicated = len(np.unique(present)) != len(present)
        for s in range(0, len(keep), chunk if not duplicated else min(chunk, 4096)):
            rows = keep[s:s + (chunk if not duplicated else min(chunk, 4096))]
            if duplicated:
                out[s:s + len(rows)] = core.project(self._dense(rows), goi, self.sf[rows].astype(np.float32), mu, sigma,
                                                    comps, mean)
                continue
            lens = ip[rows + 1] - ip[rows]
            sub_ip = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
            take = np.concatenate([np.arange(ip[r], ip[r + 1]) for r in rows]) if len(rows) else np.zeros(0, np.int64)
            out[s:s + chunk] = core.project_csr(sub_ip, ix[take], vl[take], pos, self.sf[rows].astype(np.float32),
                                                mu, sigma, comps, mean)
        h5.create_row_group(pca_group_name, [self.cells[r] for r in keep], out)
        h5.flush()
        h5.close()
        return None
