"""TEST INFRASTRUCTURE: an object with the method set of te_counter_b200._lib.Engine whose
answers come from the CPU oracle.  CPU-only tests plug it into measureTE to exercise the host
logic (packing, batching, logging, TSV writers) where no GPU exists.  Never used by the product."""
import numpy as np

from oracle import te_oracle, te_oracle_ext
from te_counter_b200 import _lib


class OracleEngine:
    def __init__(self, device=0):
        self._pinned = []
        self.n_ensg = 0
        self.options = {}

    def set_option(self, key, value):
        self.options[key] = int(value)

    def pinned(self, n, dtype):
        return np.empty(n, dtype=dtype)

    def upload_index(self, idx):
        self.idx = te_oracle.Index(idx.chrom_id, idx.L, idx.R, idx.ensg_id, idx.type_code,
                                   idx.strand_code, idx.n_ensg, idx.bucket_size)
        self.n_ensg = idx.n_ensg

    # bulk
    def bulk_begin(self, paired, qual):
        self._paired, self._qual, self._cols = paired, qual, [[] for _ in range(5)]

    def bulk_push(self, n, *cols):
        for acc, c in zip(self._cols, cols):
            acc.append(np.array(c[:n]))

    def bulk_finish(self):
        cols = [np.concatenate(c) if c else np.zeros(0, np.int64) for c in self._cols]
        st = np.zeros(_lib.BULK_NSTATS, np.int64)
        try:
            fn = te_oracle_ext.bulk_count_stranded if self.options.get("bulk_strand") else te_oracle.bulk_count
            counts, s = fn(self.idx, self._paired, self._qual, *cols)
        except te_oracle.ReferenceCrash as e:
            counts = [0] * self.n_ensg
            st[_lib.BS_CRASH_NAME if e.kind == "AttributeError" else _lib.BS_CRASH_ENHANCER] = 1
            return np.array(counts, np.int64), st
        st[_lib.BS_UNITS] = s["total_reads"] - 1
        st[_lib.BS_ASSIGNED], st[_lib.BS_LOWQ] = s["assigned"], s["lowq"]
        st[_lib.BS_BADCHROM], st[_lib.BS_QCFAIL] = s["badchrom"], s["qcfail"]
        return np.array(counts, np.int64), st

    # single cell
    def sc_begin(self, qual, strand, n_whitelist):
        self._qual, self._strand, self._cols = qual, strand, [[] for _ in range(7)]

    def sc_push(self, n, *cols):
        for acc, c in zip(self._cols, cols):
            acc.append(np.array(c[:n]))

    def sc_finalize(self, bundle_keys, maxcells, pad):
        cols = [np.concatenate(c) if c else np.zeros(0, np.int64) for c in self._cols]
        cols = [c.tolist() for c in cols]
        st = np.zeros(_lib.SC_NSTATS, np.int64)
        try:
            out = te_oracle.sc_count(self.idx, self._qual, self._strand, bundle_keys, maxcells, pad, *cols)
        except te_oracle.ReferenceCrash as e:
            if e.kind == "KeyError":
                st[_lib.SS_CRASH_STRAND] = 1
                st[_lib.SS_VALID] = 1
            st[_lib.SS_UNITS] = len(cols[0])
            self._out = ([], [], st)
            return 0, 0
        s = out["stats"]
        st[_lib.SS_UNITS] = s["total_reads"] - 1
        for k, f in ((_lib.SS_INVALID_BARCODE, "invalid_barcode"), (_lib.SS_ALREADY_SEEN, "already_seen"),
                     (_lib.SS_LOWQ, "lowq"), (_lib.SS_QCFAIL, "qcfail"), (_lib.SS_VALID, "valid"),
                     (_lib.SS_ASSIGNED, "assigned"), (_lib.SS_RAW_BARCODES, "raw_barcodes"),
                     (_lib.SS_BUNDLES, "n_bundles")):
            st[k] = s[f]
        tr = sorted(out["triples"].items())
        hits = sorted(out["cell_hits"])
        self._out = (tr, hits, st)
        return len(tr), len(hits)

    def sc_fetch(self, n_triples, n_hit):
        tr, hits, st = self._out
        return (np.array([k[0] for k, _ in tr], np.int32), np.array([k[1] for k, _ in tr], np.uint32),
                np.array([v for _, v in tr], np.int64), np.array([c for c, _ in hits], np.uint32),
                np.array([h for _, h in hits], np.int64), st)

    def sc_select(self, maxcells, n_hit):
        _, hits, _ = self._out
        order = sorted(hits, key=lambda t: (-t[1], t[0]))[:maxcells]
        return np.array([c for c, _ in order], np.uint32)

    def index_note(self):
        return ""

    def trim(self):
        pass
