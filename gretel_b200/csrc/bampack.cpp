// BAM -> packed reads on the CPU: multi-threaded BGZF inflate + one CIGAR walk per alignment.
//
// This is the producer side of the boundary (north_star keeps parsing on the CPU).  It replaces the
// reference's pysam column pileup (gretel/util.py:112-210, one Python object per read x SNP column)
// and yields exactly what that loop accumulates per read:
//   * read key / mate separation ....... util.py:149-160 (every BAM record is its own read)
//   * window ownership / start clamp ... util.py:162-176 (the union over work blocks == one pass)
//   * allele at a SNP column ........... util.py:180-190, 238 ('-' inside a deletion or ref-skip,
//                                         else the aligned base; only the first character is used)
//   * rank ............................. util.py:198  (#SNPs before the read's leftmost position)
//   * stepper filters .................. cmd.py:39,78 (samtools: UNMAP/SECONDARY/QCFAIL/DUP + orphans)
//   * reads with < 2 covered SNPs ...... dropped (util.py:230)
// Same semantics as gretel_b200/bamio.py (the dependency-free Python packer the tests compare against).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "hanselx.h"

void hx_set_error(const char *fmt, ...);

namespace {

struct Block { size_t coff, csize, uoff, usize; };

inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline int32_t rdi32(const uint8_t *p) { return (int32_t)rd32(p); }

// Splits a BGZF file into its blocks (offset/size of the raw deflate payload and of the output).
bool scan_bgzf(const std::vector<uint8_t> &f, std::vector<Block> &blocks, size_t &total) {
    size_t p = 0;
    total = 0;
    while (p + 18 <= f.size()) {
        const uint8_t *h = f.data() + p;
        if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return false;
        const size_t xlen = rd16(h + 10);
        size_t q = p + 12, xend = q + xlen;
        size_t bsize = 0;
        while (q + 4 <= xend) {
            const uint8_t *s = f.data() + q;
            const size_t slen = rd16(s + 2);
            if (s[0] == 'B' && s[1] == 'C' && slen == 2) bsize = (size_t)rd16(s + 4) + 1;
            q += 4 + slen;
        }
        if (!bsize || p + bsize > f.size()) return false;
        const size_t usize = rd32(f.data() + p + bsize - 4);
        blocks.push_back({xend, bsize - (xend - p) - 8, total, usize});
        total += usize;
        p += bsize;
    }
    return p == f.size();
}

bool inflate_block(const uint8_t *src, size_t n, uint8_t *dst, size_t m) {
    if (m == 0) return true;
    z_stream z;
    memset(&z, 0, sizeof(z));
    if (inflateInit2(&z, -15) != Z_OK) return false;
    z.next_in = const_cast<uint8_t *>(src);
    z.avail_in = (uInt)n;
    z.next_out = dst;
    z.avail_out = (uInt)m;
    const int rc = inflate(&z, Z_FINISH);
    inflateEnd(&z);
    return rc == Z_STREAM_END && z.avail_out == 0;
}

struct Out {
    std::vector<int32_t> rank;
    std::vector<int64_t> klen;
    std::vector<uint8_t> codes;
};

const uint8_t NT16_CODE[16] = {4, 0, 1, 4, 2, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4};   // "=ACMGRSVTWYHKDBN" -> A0 C1 G2 T3 else N4

// One alignment -> (rank, codes) appended to `o` when it covers at least two SNPs.
void pack_record(const uint8_t *rec, int32_t target_tid, int32_t start_pos, int32_t end_pos,
                 const int32_t *snp, int32_t n_snps, int stepper, Out &o, std::vector<uint8_t> &tmp) {
    const int32_t tid = rdi32(rec), pos = rdi32(rec + 4);
    const int l_read_name = rec[8];
    const int n_cigar = rd16(rec + 12);
    const int flag = rd16(rec + 14);
    const int32_t l_seq = rdi32(rec + 16);
    if (tid != target_tid || pos < 0) return;
    if (stepper != 2) {
        if (flag & (0x4 | 0x100 | 0x200 | 0x400)) return;
        if (stepper == 0 && (flag & 0x1) && !(flag & 0x2)) return;          // orphan rule
    }
    if (pos + 1 > end_pos) return;
    const uint8_t *cig = rec + 32 + l_read_name;
    const uint8_t *seq = cig + 4 * (size_t)n_cigar;
    int64_t qalen = 0, rlen = 0;
    for (int i = 0; i < n_cigar; ++i) {
        const uint32_t c = rd32(cig + 4 * i);
        const int op = c & 0xf;
        const int64_t ln = c >> 4;
        if (op == 0 || op == 1 || op == 7 || op == 8) qalen += ln;           // M I = X
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += ln; // M D N = X
    }
    int64_t leftmost = (int64_t)pos + 1;
    if (leftmost < start_pos) {                                               // util.py:165-171
        if ((int64_t)pos + 1 + qalen < start_pos) return;
        leftmost = start_pos;
    }
    const int32_t *sb = snp, *se = snp + n_snps;
    const int32_t rank = (int32_t)(std::lower_bound(sb, se, (int32_t)leftmost) - sb);
    const int64_t last = std::min<int64_t>((int64_t)pos + rlen, end_pos);
    const int32_t hi = (int32_t)(std::upper_bound(sb, se, (int32_t)std::min<int64_t>(last, INT32_MAX)) - sb);
    if (hi - rank < 2) return;
    tmp.clear();
    int32_t wi = rank;
    int64_t rpos = (int64_t)pos + 1, qpos = 0;
    for (int i = 0; i < n_cigar && wi < hi; ++i) {
        const uint32_t c = rd32(cig + 4 * i);
        const int op = c & 0xf;
        const int64_t ln = c >> 4;
        if (op == 0 || op == 7 || op == 8) {
            while (wi < hi && snp[wi] < rpos + ln) {
                if (snp[wi] >= rpos) {
                    const int64_t q = qpos + (snp[wi] - rpos);
                    uint8_t code = 4;
                    if (q < l_seq) {
                        const uint8_t b = seq[q >> 1];
                        code = NT16_CODE[(q & 1) ? (b & 0xf) : (b >> 4)];
                    }
                    tmp.push_back(code);
                }
                ++wi;
            }
            rpos += ln; qpos += ln;
        } else if (op == 2 || op == 3) {
            while (wi < hi && snp[wi] < rpos + ln) {
                if (snp[wi] >= rpos) tmp.push_back(5);                           // '-'
                ++wi;
            }
            rpos += ln;
        } else if (op == 1 || op == 4) {
            qpos += ln;
        }
    }
    if (tmp.size() < 2) return;
    o.rank.push_back(rank);
    o.klen.push_back((int64_t)tmp.size());
    o.codes.insert(o.codes.end(), tmp.begin(), tmp.end());
}

}  // namespace

namespace {

// Whole BAM in memory: inflated bytes, target contig id and the offset of every alignment record.
struct LoadedBam {
    std::vector<uint8_t> data;
    std::vector<size_t> recs;
    int32_t target_tid = -1;
    int32_t target_len = 0;
};

int load_bam(const char *bam_path, const char *contig, int n_threads, LoadedBam &lb) {
    FILE *fp = fopen(bam_path, "rb");
    if (!fp) { hx_set_error("cannot open %s", bam_path); return HX_E_ARG; }
    fseek(fp, 0, SEEK_END);
    const long fsz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    std::vector<uint8_t> file((size_t)fsz);
    if (fsz && fread(file.data(), 1, (size_t)fsz, fp) != (size_t)fsz) { fclose(fp); hx_set_error("short read on %s", bam_path); return HX_E_ARG; }
    fclose(fp);
    std::vector<Block> blocks;
    size_t total = 0;
    if (!scan_bgzf(file, blocks, total)) { hx_set_error("%s is not a BGZF file", bam_path); return HX_E_ARG; }
    lb.data.resize(total);
    std::atomic<size_t> next(0);
    std::atomic<bool> ok(true);
    auto work = [&]() {
        for (;;) {
            const size_t b = next.fetch_add(1);
            if (b >= blocks.size()) break;
            if (!inflate_block(file.data() + blocks[b].coff, blocks[b].csize, lb.data.data() + blocks[b].uoff, blocks[b].usize))
                ok = false;
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; ++t) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
    if (!ok) { hx_set_error("inflate failed on %s", bam_path); return HX_E_ARG; }
    const uint8_t *d = lb.data.data();
    if (total < 12 || memcmp(d, "BAM\1", 4) != 0) { hx_set_error("%s is not a BAM file", bam_path); return HX_E_ARG; }
    size_t p = 4;
    const int32_t l_text = rdi32(d + p); p += 4 + (size_t)l_text;
    const int32_t n_ref = rdi32(d + p); p += 4;
    for (int32_t i = 0; i < n_ref; ++i) {
        const int32_t l_name = rdi32(d + p); p += 4;
        const bool hit = std::string((const char *)d + p, (size_t)std::max(0, l_name - 1)) == contig;
        p += (size_t)l_name;
        if (hit) { lb.target_tid = i; lb.target_len = rdi32(d + p); }
        p += 4;
    }
    if (lb.target_tid < 0) { hx_set_error("contig %s not in %s", contig, bam_path); return HX_E_ARG; }
    while (p + 4 <= total) {
        const int32_t bs = rdi32(d + p);
        if (bs < 32 || p + 4 + (size_t)bs > total) break;
        lb.recs.push_back(p + 4);
        p += 4 + (size_t)bs;
    }
    return HX_OK;
}

}  // namespace

extern "C" {

int hx_pack_bam(const char *bam_path, const char *contig, int32_t start_pos, int32_t end_pos,
                const int32_t *snp_pos, int32_t n_snps, int stepper, int n_threads, hx_packed *out) {
    if (!bam_path || !contig || !out || n_snps < 0 || (n_snps && !snp_pos) || stepper < 0 || stepper > 2) {
        hx_set_error("hx_pack_bam: bad arguments");
        return HX_E_ARG;
    }
    memset(out, 0, sizeof(*out));
    if (n_threads < 1) n_threads = 1;
    const bool verbose = getenv("HX_PACK_VERBOSE") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!verbose) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[hx_pack_bam] %-10s %.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    };
    LoadedBam lb;
    {
        const int rc = load_bam(bam_path, contig, n_threads, lb);
        if (rc) return rc;
    }
    const std::vector<uint8_t> &data = lb.data;
    const std::vector<size_t> &recs = lb.recs;
    const int32_t target_tid = lb.target_tid;
    const size_t total = data.size();
    (void)total;
    lap("index");
    // parallel CIGAR walks over contiguous chunks of records (keeps BAM order)
    const size_t nrec = recs.size();
    const int nt = (int)std::min<size_t>((size_t)n_threads, std::max<size_t>(1, nrec / 4096));
    std::vector<Out> outs((size_t)nt);
    {
        auto work = [&](int t) {
            std::vector<uint8_t> tmp;
            const size_t a = nrec * (size_t)t / (size_t)nt, b = nrec * (size_t)(t + 1) / (size_t)nt;
            for (size_t i = a; i < b; ++i)
                pack_record(data.data() + recs[i], target_tid, start_pos, end_pos, snp_pos, n_snps, stepper, outs[(size_t)t], tmp);
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
        work(0);
        for (auto &t : th) t.join();
    }
    lap("walk");
    int64_t R = 0, C = 0;
    for (auto &o : outs) { R += (int64_t)o.rank.size(); C += (int64_t)o.codes.size(); }
    out->rank = (int32_t *)malloc(sizeof(int32_t) * (size_t)std::max<int64_t>(R, 1));
    out->off = (int64_t *)malloc(sizeof(int64_t) * (size_t)(R + 1));
    out->codes = (uint8_t *)malloc((size_t)std::max<int64_t>(C, 1));
    if (!out->rank || !out->off || !out->codes) { hx_pack_free(out); return HX_E_NOMEM; }
    int64_t r = 0, c = 0;
    out->off[0] = 0;
    for (auto &o : outs) {
        if (!o.rank.empty()) memcpy(out->rank + r, o.rank.data(), sizeof(int32_t) * o.rank.size());
        for (size_t i = 0; i < o.klen.size(); ++i) { out->off[r + 1] = out->off[r] + o.klen[i]; ++r; }
        if (!o.codes.empty()) memcpy(out->codes + c, o.codes.data(), o.codes.size());
        c += (int64_t)o.codes.size();
    }
    lap("gather");
    out->n_reads = R;
    out->n_codes = C;
    out->n_records = (int64_t)nrec;
    return HX_OK;
}

/* Per-position A,C,G,T counts over [start0, end0) of a contig from every alignment (no filter, no base
 * quality threshold): what gretel/snpper.py:30 asks pysam's count_coverage for.  out[4][end0-start0]. */
int hx_count_coverage(const char *bam_path, const char *contig, int32_t start0, int32_t end0, int n_threads,
                      uint32_t *out) {
    if (!bam_path || !contig || !out || start0 < 0 || end0 < start0) { hx_set_error("hx_count_coverage: bad arguments"); return HX_E_ARG; }
    if (n_threads < 1) n_threads = 1;
    LoadedBam lb;
    int rc = load_bam(bam_path, contig, n_threads, lb);
    if (rc) return rc;
    const int64_t len = (int64_t)end0 - start0;
    memset(out, 0, sizeof(uint32_t) * 4 * (size_t)len);
    const size_t nrec = lb.recs.size();
    const int nt = (int)std::min<size_t>((size_t)n_threads, std::max<size_t>(1, nrec / 4096));
    std::vector<std::vector<uint32_t>> part((size_t)nt);
    auto work = [&](int t) {
        std::vector<uint32_t> &c = part[(size_t)t];
        c.assign(4 * (size_t)len, 0u);
        const size_t a = nrec * (size_t)t / (size_t)nt, b = nrec * (size_t)(t + 1) / (size_t)nt;
        for (size_t i = a; i < b; ++i) {
            const uint8_t *rec = lb.data.data() + lb.recs[i];
            const int32_t tid = rdi32(rec), pos = rdi32(rec + 4);
            if (tid != lb.target_tid || pos < 0) continue;
            const int l_read_name = rec[8];
            const int n_cigar = rd16(rec + 12);
            const int32_t l_seq = rdi32(rec + 16);
            const uint8_t *cig = rec + 32 + l_read_name;
            const uint8_t *seq = cig + 4 * (size_t)n_cigar;
            int64_t rpos = pos, qpos = 0;
            for (int ci = 0; ci < n_cigar; ++ci) {
                const uint32_t cv = rd32(cig + 4 * ci);
                const int op = cv & 0xf;
                const int64_t ln = cv >> 4;
                if (op == 0 || op == 7 || op == 8) {
                    for (int64_t j = 0; j < ln; ++j) {
                        const int64_t r = rpos + j, q = qpos + j;
                        if (r < start0 || r >= end0 || q >= l_seq) continue;
                        const uint8_t bb = seq[q >> 1];
                        const uint8_t code = NT16_CODE[(q & 1) ? (bb & 0xf) : (bb >> 4)];
                        if (code < 4) c[(size_t)code * (size_t)len + (size_t)(r - start0)]++;
                    }
                    rpos += ln; qpos += ln;
                } else if (op == 2 || op == 3) {
                    rpos += ln;
                } else if (op == 1 || op == 4) {
                    qpos += ln;
                }
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto &t : th) t.join();
    for (int t = 0; t < nt; ++t)
        for (size_t i = 0; i < 4 * (size_t)len; ++i) out[i] += part[(size_t)t][i];
    return HX_OK;
}

int hx_bam_contig_length(const char *bam_path, const char *contig, int32_t *length) {
    if (!bam_path || !contig || !length) { hx_set_error("hx_bam_contig_length: bad arguments"); return HX_E_ARG; }
    LoadedBam lb;
    int rc = load_bam(bam_path, contig, 1, lb);
    if (rc) return rc;
    *length = lb.target_len;
    return HX_OK;
}

// ---- dense wire format encoder (CPU side of hx_ingest_host_dense; layout documented in wire.cu) ----------
static inline int64_t al16(int64_t x) { return (x + 15) & ~(int64_t)15; }

int hx_dense_encode(const int32_t *rank, const int64_t *off, const uint8_t *codes, int64_t n_reads, int n_threads,
                    hx_dense *out) {
    if (!out || n_reads < 0 || (n_reads > 0 && (!rank || !off || !codes))) {
        hx_set_error("hx_dense_encode: bad arguments");
        return HX_E_ARG;
    }
    memset(out, 0, sizeof(*out));
    const int64_t c0 = n_reads ? off[0] : 0, n_codes = n_reads ? off[n_reads] - c0 : 0;
    if (n_codes < 0 || n_codes >= ((int64_t)1 << 32)) {
        hx_set_error("hx_dense_encode: %lld alleles in one chunk (limit 2^32 - 1): split the reads", (long long)n_codes);
        return HX_E_ARG;
    }
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, 1 + n_reads / 65536));
    // pass 1: per-thread maxima and counts over a contiguous range of reads (and of their alleles)
    std::vector<int64_t> n_esc(nt, 0), n_exc(nt, 0), kmax(nt, 0);
    std::vector<int> bad(nt, 0);
    auto range = [&](int t, int64_t &a, int64_t &b) { a = n_reads * t / nt; b = n_reads * (t + 1) / nt; };
    auto pass1 = [&](int t) {
        int64_t a, b;
        range(t, a, b);
        for (int64_t r = a; r < b; ++r) {
            const int64_t d = (int64_t)rank[r] - (r ? (int64_t)rank[r - 1] : 0), k = off[r + 1] - off[r];
            if (d < 0 || k < 0 || k > 65535) bad[t] = d < 0 ? 1 : 2;
            if (d >= 255) n_esc[t]++;
            kmax[t] = std::max(kmax[t], k);
        }
        if (b > a && !bad[t])
            for (int64_t i = off[a]; i < off[b]; ++i) {
                if (codes[i] > 6) bad[t] = 3;
                n_exc[t] += codes[i] >= 4;
            }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(pass1, t);
        pass1(0);
        for (auto &x : th) x.join();
    }
    int64_t tot_esc = 0, tot_exc = 0, km = 0;
    for (int t = 0; t < nt; ++t) {
        if (bad[t]) {
            hx_set_error("hx_dense_encode: %s", bad[t] == 1 ? "reads are not sorted by rank"
                                               : bad[t] == 2 ? "a read covers more than 65535 SNPs (or off[] decreases)"
                                                             : "allele code > 6");
            return HX_E_ARG;
        }
        const int64_t e = n_esc[t], x = n_exc[t];
        n_esc[t] = tot_esc; n_exc[t] = tot_exc;       // exclusive prefix: where each thread writes its lists
        tot_esc += e; tot_exc += x;
        km = std::max(km, kmax[t]);
    }
    const int kb = km < 256 ? 1 : 2;
    const int64_t n_words = (n_codes + 15) / 16;
    const int64_t o_kl = al16(n_reads), o_c2 = o_kl + al16(n_reads * kb), o_ex = o_c2 + al16(n_words * 4);
    const int64_t o_ei = o_ex + al16(tot_exc * 4), o_ed = o_ei + al16(tot_esc * 8), bytes = o_ed + al16(tot_esc * 4) + 16;
    uint8_t *blob = (uint8_t *)calloc(1, (size_t)bytes);
    if (!blob) { hx_set_error("hx_dense_encode: out of memory (%lld bytes)", (long long)bytes); return HX_E_NOMEM; }
    uint32_t *exc = (uint32_t *)(blob + o_ex);
    int64_t *ei = (int64_t *)(blob + o_ei);
    int32_t *ed = (int32_t *)(blob + o_ed);
    // pass 2: a thread's alleles start at off[a]-c0, which need not be a multiple of 4: boundaries move up to the
    // next byte so that every output byte has exactly one writer
    auto pass2 = [&](int t) {
        int64_t a, b;
        range(t, a, b);
        int64_t ne = n_esc[t];
        for (int64_t r = a; r < b; ++r) {
            const int64_t d = (int64_t)rank[r] - (r ? (int64_t)rank[r - 1] : 0), k = off[r + 1] - off[r];
            blob[r] = (uint8_t)std::min<int64_t>(d, 255);
            if (d >= 255) { ei[ne] = r; ed[ne] = (int32_t)d; ++ne; }
            if (kb == 1) blob[o_kl + r] = (uint8_t)k; else ((uint16_t *)(blob + o_kl))[r] = (uint16_t)k;
        }
        if (b <= a) return;
        // alleles [lo, hi) of the stream, rounded so that every output byte has exactly one writer
        int64_t lo = off[a] - c0, hi = off[b] - c0;
        lo = t == 0 ? 0 : (lo + 3) & ~(int64_t)3;
        hi = t == nt - 1 ? n_codes : (hi + 3) & ~(int64_t)3;
        hi = std::min(hi, n_codes);
        uint8_t *c2 = blob + o_c2;
        for (int64_t i = lo; i < hi; i += 4) {
            uint8_t byte = 0;
            const int64_t m = std::min<int64_t>(4, n_codes - i);
            for (int64_t j = 0; j < m; ++j) {
                const uint8_t c = codes[c0 + i + j];
                byte |= (uint8_t)((c >= 4 ? c - 4 : c) << (2 * j));
            }
            c2[i >> 2] = byte;
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(pass2, t);
        pass2(0);
        for (auto &x : th) x.join();
    }
    // exception positions: per-thread slots were sized by read ranges, so fill them by read ranges too
    auto pass3 = [&](int t) {
        int64_t a, b;
        range(t, a, b);
        if (b <= a) return;
        int64_t nx = n_exc[t];
        for (int64_t i = off[a] - c0; i < off[b] - c0; ++i)
            if (codes[c0 + i] >= 4) exc[nx++] = (uint32_t)i;
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(pass3, t);
        pass3(0);
        for (auto &x : th) x.join();
    }
    out->blob = blob; out->blob_bytes = bytes; out->n_reads = n_reads; out->n_codes = n_codes;
    out->n_exc = tot_exc; out->n_esc = tot_esc; out->klen_bytes = kb;
    out->o_klen = o_kl; out->o_codes2 = o_c2; out->o_exc = o_ex; out->o_esc_idx = o_ei; out->o_esc_delta = o_ed;
    return HX_OK;
}

void hx_dense_free(hx_dense *d) {
    if (!d) return;
    free(d->blob);
    memset(d, 0, sizeof(*d));
}

void hx_pack_free(hx_packed *p) {
    if (!p) return;
    free(p->rank); free(p->off); free(p->codes);
    memset(p, 0, sizeof(*p));
}

}  // extern "C"
