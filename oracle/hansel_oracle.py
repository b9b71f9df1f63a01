"""CPU oracle for the Gretel/Hansel hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product (``gretel_b200``)
never imports it and has no CPU fallback.

What it restates (citations are ``path:line`` under /root/reference):

* ``ingest_packed``        <- gretel/util.py:226-286  (pair expansion + sentinel rules)
* ``load_totals``          <- gretel/util.py:329-333  (n_slices, n_crumbs, L)
* ``reweight_hansel_from_path`` <- gretel/gretel.py:79-98 (literal loop nest, incl. the
  double hit of adjacent pairs and the never-touched (p,N) pairs)
* ``generate_path``        <- gretel/gretel.py:136-189 (first-max walk, log10 sums, min marginal)
* ``recover``              <- gretel/cmd.py:148-179   (ratio clamp, loop, PATHS bookkeeping)
* ``OracleHansel``         <- the surface of the un-vendored dependency ``hanselx==0.0.92``
  (setup.py:8) as used at util.py:83,266-286, gretel.py:84,96,155,182,186,
  cmd.py:79,82,86,127,201 and tests/test_test.py:41-52.

PARITY STATUS
-------------
* Ingestion (init_matrix/add_observation/get_observation, n_slices, n_crumbs):
  PINNED by the reference's own test (tests/test_test.py:33-52) - see
  tests/test_oracle_golden.py which decodes tests/golden/ref_test.bam + ref_test.vcf.gz
  (byte copies of the reference's binary *fixtures*, not source) and checks all 12
  known answers for n_threads-independent ingestion.
* Recovery arithmetic (get_counts_at, get_marginal_of_at, get_edge_weights_at,
  reweight_observation): **PARITY UNPINNED**.  hanselx is not in /root/reference, is not
  installed and cannot be fetched (no network); no reference test pins these.  What
  is below is the call-site contract (hard) plus the published/recollected upstream
  algorithm; every uncertain choice is a named switch on ``OracleHansel``:
    - ``v_site``: Laplace denominator uses the number of valid symbols seen at the
      *from* site ("from", default - gretel.py:10 TODO names both options) or at the
      *to* site ("to").
    - ``candidates_skip_unsymbols``: N and _ are never offered as branches (default True).
    - storage dtype float32 (fork-shared ctypes.c_float array upstream), arithmetic float64.
  One documented deviation: a lookback term whose Laplace denominator is 0 (only
  possible looking back to the start sentinel from snp>=2 with v_site="from") is
  dropped instead of raising ZeroDivisionError.

Everything here is deliberately the slow, literal, dense (7,7,N+2,N+2) formulation.
The fast C restatement used for full-size checks lives in oracle/hansel_oracle.c and
is itself validated against this file in tests/.
"""
from __future__ import annotations

import sys
from math import ceil, log10

import numpy as np

SYMBOLS = ['A', 'C', 'G', 'T', 'N', '-', '_']   # util.py:83 (order defines codes 0..6)
UNSYMBOLS = ['N', '_']                          # util.py:83
CODE = {s: i for i, s in enumerate(SYMBOLS)}


class OracleHansel:
    """Dense, literal Hansel (see module docstring for what is pinned and what is not)."""

    def __init__(self, symbols, unsymbols, n_snps, dtype=np.float32,
                 v_site="from", candidates_skip_unsymbols=True):
        self.symbols = list(symbols)
        self.unsymbols = list(unsymbols)
        self.symbols_d = {s: s for s in self.symbols}      # gretel.py:138, cmd.py:201
        self._idx = {s: i for i, s in enumerate(self.symbols)}
        self.n_snps = int(n_snps)
        self.m = np.zeros((len(symbols), len(symbols), n_snps + 2, n_snps + 2), dtype=dtype)
        self.n_slices = 0
        self.n_crumbs = 0
        self.L = 1
        self.v_site = v_site
        self.candidates_skip_unsymbols = candidates_skip_unsymbols

    # util.py:83
    @classmethod
    def init_matrix(cls, symbols, unsymbols, n_snps, **kw):
        return cls(symbols, unsymbols, n_snps, **kw)

    def copy(self):                                         # cmd.py:79
        o = OracleHansel(self.symbols, self.unsymbols, self.n_snps, dtype=self.m.dtype,
                         v_site=self.v_site,
                         candidates_skip_unsymbols=self.candidates_skip_unsymbols)
        o.m[...] = self.m
        o.n_slices, o.n_crumbs, o.L = self.n_slices, self.n_crumbs, self.L
        return o

    # util.py:266-286
    def add_observation(self, a, b, i, j):
        self.m[self._idx[a], self._idx[b], i, j] += 1

    # tests/test_test.py:41-52
    def get_observation(self, a, b, i, j):
        return float(self.m[self._idx[a], self._idx[b], i, j])

    # gretel.py:84,96  (UNPINNED: new = old - ratio*old in float64, stored as float32)
    def reweight_observation(self, a, b, i, j, ratio):
        old = float(self.m[self._idx[a], self._idx[b], i, j])
        new = old - (float(ratio) * old)
        self.m[self._idx[a], self._idx[b], i, j] = new
        return old - new

    # gretel.py:72 (dead code upstream; kept for surface completeness)
    def reweight_matrix(self, ratio):
        m64 = self.m.astype(np.float64)
        self.m[...] = m64 - float(ratio) * m64

    # cmd.py:86-92,127-143  (UNPINNED)
    def get_counts_at(self, at_pos):
        out = {}
        total = 0.0
        for s in self.symbols:
            c = 0.0
            row = self.m[self._idx[s], :, at_pos, at_pos + 1]
            for b in range(len(self.symbols)):          # fixed order, float64 accumulate
                c += float(row[b])
            if c > 0:
                out[s] = c
                total += c
        out["total"] = total
        return out

    # gretel.py:182,186  (UNPINNED)
    def get_marginal_of_at(self, of_symbol, at_pos):
        counts = self.get_counts_at(at_pos)
        if counts["total"] == 0:
            return 0.0
        return counts.get(of_symbol, 0.0) / counts["total"]

    def _valid_symbols_seen(self, at_pos):
        counts = self.get_counts_at(at_pos)
        return sum(1 for s in self.symbols
                   if s not in self.unsymbols and counts.get(s, 0.0) > 0)

    def get_spanning_support(self, symbol_to, pos_from, pos_to):
        t = 0.0
        col = self.m[:, self._idx[symbol_to], pos_from, pos_to]
        for a in range(len(self.symbols)):
            t += float(col[a])
        return t

    # P(symbol_from at pos_from | symbol_to at pos_to), Laplace smoothed  (UNPINNED)
    def get_conditional_of_at(self, symbol_from, symbol_to, pos_from, pos_to):
        obs = self.get_observation(symbol_from, symbol_to, pos_from, pos_to)
        tot = self.get_spanning_support(symbol_to, pos_from, pos_to)
        v = self._valid_symbols_seen(pos_from if self.v_site == "from" else pos_to)
        den = v + tot
        if den == 0:
            return None
        return (1.0 + obs) / den

    # gretel.py:155  (UNPINNED)
    def get_edge_weights_at(self, snp, current_path, debug=False):
        counts = self.get_counts_at(snp)
        total = counts["total"]
        branches = {}
        tot_w = 0.0
        for s in self.symbols:
            if self.candidates_skip_unsymbols and s in self.unsymbols:
                continue
            c = counts.get(s, 0.0)
            if not c > 0:
                continue
            lw = log10(c / total)
            for l in range(1, min(self.L, snp) + 1):
                pos_from = snp - l
                p = self.get_conditional_of_at(current_path[pos_from], s, pos_from, snp)
                if p is None:
                    continue
                lw += log10(p)
            w = 10.0 ** lw
            branches[s] = w
            tot_w += w
        if tot_w > 0:
            for s in list(branches):
                branches[s] = branches[s] / tot_w
        branches["total"] = tot_w
        return branches

    def save_hansel_dump(self, path):                       # cmd.py:82
        np.save(path, self.m)


# --------------------------------------------------------------------------- ingestion
def ingest_packed(hansel, rank, off, codes, n_snps, use_end_sentinels=False):
    """util.py:226-286 over packed reads.  Returns (slices, crumbs, covered_snps)."""
    slices = crumbs = covered = 0
    N = n_snps
    for r in range(len(rank)):
        seq = [SYMBOLS[c] for c in codes[off[r]:off[r + 1]]]
        if not len(seq) > 1:                                # util.py:230
            continue
        slices += 1
        rk = int(rank[r])
        support_len = len(seq)
        support_seq = "".join(seq)
        covered += len(support_seq.replace("N", "").replace("_", ""))   # util.py:239
        for i in range(0, support_len):
            snp_a = support_seq[i]
            for j in range(i + 1, support_len):
                snp_b = support_seq[j]
                if snp_a in ['_', 'N']:                     # util.py:258
                    continue
                if i == 0 and j == 1 and rk == 0:           # util.py:262
                    hansel.add_observation('_', snp_a, 0, 1)
                    hansel.add_observation(snp_a, snp_b, 1, 2)
                    crumbs += 1
                elif (j + rk + 1) == N and abs(i - j) == 1:  # util.py:271
                    hansel.add_observation(snp_a, snp_b, N - 1, N)
                    hansel.add_observation(snp_b, '_', N, N + 1)
                    crumbs += 1
                else:                                        # util.py:279
                    hansel.add_observation(snp_a, snp_b, i + rk + 1, j + rk + 1)
                    crumbs += 1
                    if use_end_sentinels:
                        if j == (support_len - 1) and abs(i - j) == 1:
                            hansel.add_observation(snp_b, '_', j + rk + 1, j + rk + 2)
    return slices, crumbs, covered


def load_from_packed(rank, off, codes, n_snps, **hansel_kw):
    """util.py:83 + 226-286 + 329-333."""
    h = OracleHansel.init_matrix(SYMBOLS, UNSYMBOLS, n_snps, **hansel_kw)
    slices, crumbs, covered = ingest_packed(h, rank, off, codes, n_snps)
    h.n_slices = slices
    h.n_crumbs = crumbs
    h.L = int(ceil(float(covered) / slices))                # util.py:333
    return h


# --------------------------------------------------------------------------- recovery
def reweight_hansel_from_path(hansel, path, ratio):
    """gretel.py:79-98, literal."""
    size = 0
    for i in range(0, len(path)):
        for j in range(0, i + 1 + 1):
            if i >= len(path) - 1:
                size += hansel.reweight_observation(path[i], path[j], i, i + 1, ratio)
                break
            else:
                if j < i:
                    t_i = j
                    t_j = i
                else:
                    t_i = i
                    t_j = j
                size += hansel.reweight_observation(path[t_i], path[t_j], t_i, t_j, ratio)
    return size


def generate_path(n_snps, hansel, original_hansel):
    """gretel.py:136-189, literal (debug hooks dropped)."""
    running_prob = 0.0
    running_prob_uw = 0.0
    current_path = [hansel.symbols_d['_']]
    marginals = []
    for snp in range(1, n_snps + 1):
        curr_branches = hansel.get_edge_weights_at(snp, current_path)
        next_v = 0.0
        next_m = None
        for symbol in curr_branches:
            if str(symbol) == "total":
                continue
            if next_m is None:
                next_v = curr_branches[symbol]
                next_m = symbol
            elif curr_branches[symbol] > next_v:
                next_v = curr_branches[symbol]
                next_m = symbol
        if next_m is None:
            return None, None, None
        selected_edge_weight = hansel.get_marginal_of_at(next_m, snp)
        marginals.append(selected_edge_weight)
        running_prob += log10(selected_edge_weight)
        running_prob_uw += log10(original_hansel.get_marginal_of_at(next_m, snp))
        current_path.append(next_m)
    return current_path, {"hp_original": running_prob_uw, "hp_current": running_prob}, min(marginals)


def recover(hansel, n_snps, max_paths=100, min_remove=0.01):
    """cmd.py:79,148-179: copy, loop, ratio clamp, PATHS bookkeeping.

    Returns list of per-iteration dicts (path string, hp_current, hp_original, min
    marginal, ratio used, removed) and the PATHS dict keyed like cmd.py:164."""
    original = hansel.copy()
    iters = []
    PATHS = {}
    for i in range(max_paths):
        path, prob, init_min = generate_path(n_snps, hansel, original)
        if path is None:
            break
        ratio = init_min
        if ratio < min_remove:
            ratio = min_remove
        mag = reweight_hansel_from_path(hansel, path, ratio)
        key = "".join(str(x) for x in path)
        if key not in PATHS:
            PATHS[key] = {"hp_current": [], "hp_original": [], "i": [], "i_0": i, "n": 0,
                          "magnitude": 0, "hansel_path": path}
        PATHS[key]["n"] += 1
        PATHS[key]["i"].append(i)
        PATHS[key]["magnitude"] += mag
        PATHS[key]["hp_current"].append(prob["hp_current"])
        PATHS[key]["hp_original"].append(prob["hp_original"])
        iters.append({"path": key, "hp_current": prob["hp_current"],
                      "hp_original": prob["hp_original"], "min_marginal": init_min,
                      "ratio": ratio, "removed": mag})
    return iters, PATHS


def gap_check(hansel, n_snps):
    """cmd.py:85-92: list of sites i in 0..N whose counts total is 0."""
    return [i for i in range(0, n_snps + 1) if hansel.get_counts_at(i).get("total", 0) == 0]


def band_of(hansel, W):
    """Dense -> banded float32 [N+2][W][7][7] (cell (pi,pj) at [pj][pj-pi-1]); test helper."""
    N = hansel.n_snps
    out = np.zeros((N + 2, W, 7, 7), dtype=np.float32)
    for pj in range(1, N + 2):
        for d in range(1, W + 1):
            pi = pj - d
            if pi < 0:
                break
            out[pj, d - 1] = hansel.m[:, :, pi, pj]
    return out


def out_of_band_mass(hansel, W):
    """Sum of |cells| outside 1 <= pj-pi <= W; test helper."""
    N = hansel.n_snps
    tot = float(np.abs(hansel.m).sum())
    return tot - float(np.abs(band_of(hansel, W)).sum())
