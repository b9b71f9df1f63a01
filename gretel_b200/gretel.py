"""Recovery entry points with the reference's names, arguments and return values.

Mirrors gretel/gretel.py:13 (reweight_hansel_from_path) and :102 (generate_path); the
whole per-site walk and the O(N*W) reweight run on the GPU with no host round trip
per site.  ``recover`` is the driver loop of gretel/cmd.py:148-179.
"""
from __future__ import annotations

import sys


def reweight_hansel_from_path(hansel, path, ratio):
    """gretel/gretel.py:13-98 -> sum of removed observations (float)."""
    size = hansel.reweight_path_codes(hansel.encode_path(path), ratio)
    sys.stderr.write("[RWGT] Ratio %.3f, Removed %.1f\n" % (ratio, size))
    return size


def generate_path(n_snps, hansel, original_hansel, debug_hpos=None):
    """gretel/gretel.py:102-189 -> (path, {"hp_original","hp_current"}, min_marginal)
    or (None, None, None) when a hole is found (gretel.py:176-180)."""
    if n_snps != hansel.n_snps:
        raise ValueError("n_snps does not match the Hansel matrix")
    sys.stderr.write("[NOTE] *Establishing next path\n")
    res = hansel.generate_path_codes(original_hansel)
    if res[0] is None:
        snp = res[1]
        sys.stderr.write('''[NOTE] Unable to select next branch from SNP %d to %d
       By design, Gretel will attempt to recover haplotypes until a hole in the graph has been found.
       Recovery will intentionally terminate now.\n''' % (snp - 1, snp))
        return None, None, None
    codes, hp_cur, hp_orig, min_marg = res
    if debug_hpos:
        path = hansel.decode_path(codes)
        for snp in debug_hpos:
            if 1 <= snp <= n_snps:
                print(hansel.get_edge_weights_at(snp, path))
    return hansel.decode_path(codes), {"hp_original": hp_orig, "hp_current": hp_cur}, min_marg


def recover(hansel, n_snps, max_paths=100, min_remove=0.01, original_hansel=None, resident=True):
    """gretel/cmd.py:79,148-179: copy, up to ``max_paths`` x (generate, clamp ratio,
    reweight), PATHS bookkeeping.  ``resident=True`` keeps the loop on the device
    (hx_recover); ``False`` calls generate_path / reweight_hansel_from_path per haplotype
    exactly like cmd.py does.  Returns (iterations, PATHS)."""
    original = original_hansel if original_hansel is not None else hansel.copy()
    iters = []
    if resident:
        codes, stats = hansel.recover_codes(original, max_paths, min_remove)
        for c, s in zip(codes, stats):
            iters.append({"hansel_path": hansel.decode_path(c), "hp_current": float(s[0]),
                          "hp_original": float(s[1]), "min_marginal": float(s[2]), "ratio": float(s[3]),
                          "removed": float(s[4])})
    else:
        for _ in range(max_paths):
            path, prob, init_min = generate_path(n_snps, hansel, original)
            if path is None:
                break
            ratio = init_min
            if ratio < min_remove:
                sys.stderr.write("[RWGT] Ratio %.10f too small, adjusting to %.3f\n" % (ratio, min_remove))
                ratio = min_remove
            mag = reweight_hansel_from_path(hansel, path, ratio)
            iters.append({"hansel_path": path, "hp_current": prob["hp_current"],
                          "hp_original": prob["hp_original"], "min_marginal": init_min, "ratio": ratio,
                          "removed": mag})
    PATHS = {}
    for i, it in enumerate(iters):
        key = "".join(str(x) for x in it["hansel_path"])
        it["path"] = key
        if key not in PATHS:
            PATHS[key] = {"hp_current": [], "hp_original": [], "i": [], "i_0": i, "n": 0, "magnitude": 0,
                          "hansel_path": it["hansel_path"]}
        PATHS[key]["n"] += 1
        PATHS[key]["i"].append(i)
        PATHS[key]["magnitude"] += it["removed"]
        PATHS[key]["hp_current"].append(it["hp_current"])
        PATHS[key]["hp_original"].append(it["hp_original"])
    return iters, PATHS


def gap_check(hansel, n_snps):
    """gretel/cmd.py:85-92: sites whose pairwise evidence total is zero."""
    return [i for i in range(0, n_snps + 1) if hansel.get_counts_at(i).get("total", 0) == 0]
