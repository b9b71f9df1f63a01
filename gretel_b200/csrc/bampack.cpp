// BAM -> packed reads on the CPU: multi-threaded BGZF inflate + one CIGAR walk per alignment.
//
// This is the producer side of the boundary (north_star keeps parsing on the CPU).  It replaces the
// reference's pysam column pileup (gretel/util.py:112-210, one Python object per read x SNP column)
// and yields exactly what that loop accumulates per read:
//   * read key / mate separation ....... util.py:149-160 (every BAM record is its own read; records that share
//                                         query name + flag + mate would be merged upstream: deliberate deviation)
//   * pysam's pileup depth cap ......... max_depth = 8000 reads buffered by bam_plp (optional, hx_pack_bam_ex)
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
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <future>
#include <memory>
#include <queue>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "hanselx.h"
#include "hx_inflate.h"

void hx_set_error(const char *fmt, ...);
int hx_launch_coverage(const int32_t *d_seg_start, const int64_t *d_seg_nib, const int32_t *d_seg_len,
                       const uint8_t *d_seq4, int64_t n_seg, int32_t start0, int32_t len, uint32_t *d_counts,
                       cudaStream_t stream);

namespace {

inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline int32_t rdi32(const uint8_t *p) { return (int32_t)rd32(p); }

// zlib's inflate: the verdict on whatever the fast decoder declines
bool inflate_block_zlib(const uint8_t *src, size_t n, uint8_t *dst, size_t m) {
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

// one BGZF block: src[0..n) is the raw DEFLATE payload, followed in the file by the frame's 8-byte trailer (CRC32,
// ISIZE) - which is what lets the decoder refill with whole 8-byte loads up to the payload's last byte
bool inflate_block(const uint8_t *src, size_t n, uint8_t *dst, size_t m) {
    if (m == 0) return true;
    static const bool use_zlib = getenv("HX_INFLATE") && !strcmp(getenv("HX_INFLATE"), "zlib");
    if (!use_zlib && hxz::inflate_raw(src, n, n + 8, dst, m)) return true;
    return inflate_block_zlib(src, n, dst, m);
}

struct Out {
    std::vector<int32_t> rank;
    std::vector<int64_t> klen;
    std::vector<uint8_t> codes;
};

const uint8_t NT16_CODE[16] = {4, 0, 1, 4, 2, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4};   // "=ACMGRSVTWYHKDBN" -> A0 C1 G2 T3 else N4

// One alignment record (the bytes behind its block_size field), bounds-checked against block_size.
struct Rec {
    const uint8_t *p;
    uint32_t size;
    int32_t tid, pos, l_seq;
    int flag, n_cigar;
    const uint8_t *cig, *seq;       // CIGAR ops (possibly the CG:B,I tag of a long read) and the 4-bit bases
    bool ok;                        // false: the fixed fields do not fit block_size (truncated / corrupt record)
};

// Long CIGARs (> 65535 operations, ONT reads) are stored in the CG:B,I tag; the CIGAR field then holds the
// placeholder <l_seq>S<ref_len>N (SAM spec 4.2.2).
inline void find_cg_tag(Rec &r, const uint8_t *aux, const uint8_t *end) {
    while (aux + 3 <= end) {
        const uint8_t t0 = aux[0], t1 = aux[1], ty = aux[2];
        aux += 3;
        size_t sz = 0;
        switch (ty) {
            case 'A': case 'c': case 'C': sz = 1; break;
            case 's': case 'S': sz = 2; break;
            case 'i': case 'I': case 'f': sz = 4; break;
            case 'Z': case 'H': { const uint8_t *q = aux; while (q < end && *q) ++q; sz = (size_t)(q - aux) + 1; break; }
            case 'B': {
                if (aux + 5 > end) return;
                const uint8_t sub = aux[0];
                const uint32_t cnt = rd32(aux + 1);
                const size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
                if (t0 == 'C' && t1 == 'G' && sub == 'I' && aux + 5 + (size_t)cnt * 4 <= end) {
                    r.cig = aux + 5;
                    r.n_cigar = (int)cnt;
                    return;
                }
                sz = 5 + (size_t)cnt * es;
                break;
            }
            default: return;        // unknown type: stop rather than misparse
        }
        if (aux + sz > end) return;
        aux += sz;
    }
}

inline Rec parse_rec(const uint8_t *rec, uint32_t size) {
    Rec r;
    r.p = rec; r.size = size;
    r.tid = rdi32(rec); r.pos = rdi32(rec + 4);
    const int l_read_name = rec[8];
    r.n_cigar = rd16(rec + 12);
    r.flag = rd16(rec + 14);
    r.l_seq = rdi32(rec + 16);
    r.cig = rec + 32 + l_read_name;
    r.seq = r.cig + 4 * (size_t)r.n_cigar;
    const uint64_t need = 32ull + (uint64_t)l_read_name + 4ull * (uint64_t)r.n_cigar +
                          (r.l_seq > 0 ? ((uint64_t)r.l_seq + 1) / 2 + (uint64_t)r.l_seq : 0ull);
    r.ok = r.l_seq >= 0 && need <= size;
    if (r.ok && r.n_cigar == 2) {
        const uint32_t c0 = rd32(r.cig), c1 = rd32(r.cig + 4);
        if ((c0 & 0xf) == 4 && (int32_t)(c0 >> 4) == r.l_seq && (c1 & 0xf) == 3)
            find_cg_tag(r, rec + need, rec + size);
    }
    return r;
}

inline bool passes_stepper(int flag, int stepper) {            // gretel/cmd.py:39,78
    if (stepper == 2) return true;
    if (flag & (0x4 | 0x100 | 0x200 | 0x400)) return false;
    if (stepper == 0 && (flag & 0x1) && !(flag & 0x2)) return false;          // orphan rule
    return true;
}

// reference span of the alignment (M D N = X)
inline int64_t ref_len_of(const Rec &r) {
    int64_t rlen = 0;
    for (int i = 0; i < r.n_cigar; ++i) {
        const uint32_t c = rd32(r.cig + 4 * (size_t)i);
        const int op = c & 0xf;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += c >> 4;
    }
    return rlen;
}

// One alignment -> (rank, codes) appended to `o` when it covers at least two SNPs.
void pack_record(const Rec &r, int32_t target_tid, int32_t start_pos, int32_t end_pos,
                 const int32_t *snp, int32_t n_snps, int stepper, Out &o, std::vector<uint8_t> &tmp) {
    const int32_t pos = r.pos, l_seq = r.l_seq;
    const int n_cigar = r.n_cigar;
    if (!r.ok || r.tid != target_tid || pos < 0) return;
    if (!passes_stepper(r.flag, stepper)) return;
    if (pos + 1 > end_pos) return;
    const uint8_t *cig = r.cig, *seq = r.seq;
    int64_t qalen = 0, rlen = 0;
    for (int i = 0; i < n_cigar; ++i) {
        const uint32_t c = rd32(cig + 4 * (size_t)i);
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
        const uint32_t c = rd32(cig + 4 * (size_t)i);
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

// ---- streaming BAM reader ----------------------------------------------------------------------------------
// The file is read in "waves": a few tens of megabytes of BGZF blocks are inflated in parallel into one buffer,
// the record boundaries of the wave are found (a serial but trivial chain of block_size fields), the records
// are handed to the caller in parallel ranges, and the bytes of a record cut by the wave's end are carried
// into the next wave.  Memory stays bounded by the wave size whatever the size of the BAM, and reading stops
// as soon as a coordinate-sorted file has passed the target region.
struct RecRef { size_t off; uint32_t size; };

// one wave: the inflated bytes (behind the bytes carried over from the previous wave) and its records
struct Wave {
    std::unique_ptr<uint8_t[]> ubuf;
    size_t ucap = 0, ulen = 0;
    std::vector<RecRef> recs;
};

struct BamStream {
    int fd = -1;
    const uint8_t *fmap = nullptr;      // the whole file, mapped: the inflate threads read the page cache directly
    size_t fsize = 0, fpos = 0;
    std::string path;
    int n_threads = 1;
    // header
    bool header_done = false, sorted = false;
    int32_t target_tid = -1, target_len = 0;
    std::string contig;
    // wave state
    Wave own;                           // the wave of the single-buffer interface below
    std::unique_ptr<uint8_t[]> &ubuf = own.ubuf;
    std::vector<RecRef> &recs = own.recs;
    std::vector<uint8_t> carry;         // unconsumed tail of the previous wave
    bool eof = false;
    double t_read = 0, t_inflate = 0, t_scan = 0;

    ~BamStream() {
        if (fmap && fsize) munmap(const_cast<uint8_t *>(fmap), fsize);
        if (fd >= 0) close(fd);
    }

    int open(const char *bam_path, const char *ctg, int nt) {
        path = bam_path; contig = ctg; n_threads = std::max(1, nt);
        fd = ::open(bam_path, O_RDONLY);
        if (fd < 0) { hx_set_error("cannot open %s", bam_path); return HX_E_ARG; }
        struct stat sb;
        if (fstat(fd, &sb) != 0) { hx_set_error("cannot stat %s", bam_path); return HX_E_ARG; }
        fsize = (size_t)sb.st_size;
        if (fsize) {
            void *m = mmap(nullptr, fsize, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) { hx_set_error("cannot map %s", bam_path); return HX_E_ARG; }
            fmap = static_cast<const uint8_t *>(m);
            madvise(m, fsize, MADV_SEQUENTIAL);
        }
        return HX_OK;
    }

    // parses the BAM header at the front of ubuf; returns bytes consumed, 0 if more data is needed, -1 on error
    long parse_header(const Wave &w) {
        const uint8_t *d = w.ubuf.get();
        const size_t n = w.ulen;
        if (n < 12) return 0;
        if (memcmp(d, "BAM\1", 4) != 0) { hx_set_error("%s is not a BAM file", path.c_str()); return -1; }
        size_t p = 4;
        const int32_t l_text = rdi32(d + p);
        if (l_text < 0) { hx_set_error("%s: corrupt BAM header", path.c_str()); return -1; }
        if (p + 4 + (size_t)l_text + 4 > n) return 0;
        {
            const std::string text((const char *)d + p + 4, (size_t)l_text);
            const size_t hd = text.find("@HD");
            if (hd != std::string::npos) {
                const size_t eol = text.find('\n', hd);
                sorted = text.substr(hd, eol == std::string::npos ? std::string::npos : eol - hd).find("SO:coordinate") != std::string::npos;
            }
        }
        p += 4 + (size_t)l_text;
        const int32_t n_ref = rdi32(d + p); p += 4;
        if (n_ref < 0) { hx_set_error("%s: corrupt BAM header", path.c_str()); return -1; }
        for (int32_t i = 0; i < n_ref; ++i) {
            if (p + 4 > n) return 0;
            const int32_t l_name = rdi32(d + p); p += 4;
            if (l_name < 1) { hx_set_error("%s: corrupt reference name", path.c_str()); return -1; }
            if (p + (size_t)l_name + 4 > n) return 0;
            const bool hit = std::string((const char *)d + p, (size_t)l_name - 1) == contig;
            p += (size_t)l_name;
            if (hit) { target_tid = i; target_len = rdi32(d + p); }
            p += 4;
        }
        if (target_tid < 0) { hx_set_error("contig %s not in %s", contig.c_str(), path.c_str()); return -1; }
        header_done = true;
        return (long)p;
    }

    // Reads, inflates and indexes the next wave.  Returns HX_OK with recs filled (possibly empty), or an error;
    // `eof` is set once the file is exhausted.
    int next_wave(size_t wave_cbytes) { return next_wave(own, wave_cbytes); }

    int next_wave(Wave &w, size_t wave_cbytes) {
        using clk = std::chrono::steady_clock;
        std::unique_ptr<uint8_t[]> &ubuf = w.ubuf;
        std::vector<RecRef> &recs = w.recs;
        size_t &ucap = w.ucap, &ulen = w.ulen;
        recs.clear();
        if (eof) return HX_OK;
        auto t0 = clk::now();
        // compressed bytes: the whole BGZF blocks that start within the next wave_cbytes of the file
        if (fpos >= fsize) { eof = true; if (!carry.empty()) { hx_set_error("%s: truncated BAM (partial record at EOF)", path.c_str()); return HX_E_ARG; } return HX_OK; }
        const uint8_t *cbase = fmap + fpos;
        const size_t avail = fsize - fpos;
        struct Blk { size_t coff, csize, uoff, usize; };
        std::vector<Blk> blocks;
        size_t p = 0, total = 0;
        while (p < wave_cbytes && p + 18 <= avail) {
            const uint8_t *h = cbase + p;
            if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) { hx_set_error("%s is not a BGZF file", path.c_str()); return HX_E_ARG; }
            const size_t xlen = rd16(h + 10);
            size_t q = p + 12;
            const size_t xend = q + xlen;
            if (xend > avail) { hx_set_error("%s: truncated BGZF block at EOF", path.c_str()); return HX_E_ARG; }
            size_t bsize = 0;
            while (q + 4 <= xend) {
                const uint8_t *sf = cbase + q;
                const size_t slen = rd16(sf + 2);
                if (sf[0] == 'B' && sf[1] == 'C' && slen == 2 && q + 6 <= xend) bsize = (size_t)rd16(sf + 4) + 1;
                q += 4 + slen;
            }
            if (!bsize || bsize < (xend - p) + 8) { hx_set_error("%s: corrupt BGZF block", path.c_str()); return HX_E_ARG; }
            if (p + bsize > avail) { hx_set_error("%s: truncated BGZF block at EOF", path.c_str()); return HX_E_ARG; }
            const size_t usize = rd32(cbase + p + bsize - 4);
            if (usize > 65536) { hx_set_error("%s: corrupt BGZF block (ISIZE)", path.c_str()); return HX_E_ARG; }
            blocks.push_back({xend, bsize - (xend - p) - 8, total, usize});
            total += usize;
            p += bsize;
        }
        if (p == 0) { hx_set_error("%s: trailing bytes that are not a BGZF block", path.c_str()); return HX_E_ARG; }
        fpos += p;
        if (fpos >= fsize) eof = true;
        auto t1 = clk::now();
        // inflate behind the carried bytes
        const size_t need = carry.size() + total;
        if (need > ucap) { ucap = need + need / 4 + 65536; ubuf.reset(new uint8_t[ucap]); }
        if (!carry.empty()) memcpy(ubuf.get(), carry.data(), carry.size());
        const size_t base = carry.size();
        ulen = need;
        // The records of the wave are found by following the chain of block_size fields - serial, one cache miss per
        // record (20 ns).  htslib never splits a record between BGZF blocks unless it is larger than a block
        // (bam_write1 flushes first), so in practice a block starts a record: the thread that has just inflated a
        // group of consecutive blocks walks the chain through them speculatively from the group's first byte while
        // they are still in its cache; the serial pass further down only checks that the real chain arrives exactly
        // where a group started - where it does not (records cut by block ends), it walks that stretch itself.
        struct Stretch { size_t start, end, exit; bool ok, wave_end; std::vector<RecRef> recs; };
        std::vector<Stretch> st;
        {
            const uint8_t *d = ubuf.get();
            const size_t gb = std::min<size_t>(16, std::max<size_t>(1, blocks.size() / (4 * (size_t)n_threads)));
            const size_t n_groups = (blocks.size() + gb - 1) / gb;
            st.resize(n_groups);
            for (size_t g = 0; g < n_groups; ++g) {
                const size_t b1 = (g + 1) * gb;
                st[g].start = base + blocks[g * gb].uoff;
                st[g].end = b1 < blocks.size() ? base + blocks[b1].uoff : ulen;
                st[g].exit = st[g].start;
                st[g].ok = false;
                st[g].wave_end = false;
            }
            auto walk = [&](Stretch &x) {
                size_t p = x.start;
                x.recs.reserve((x.end - x.start) / 200 + 16);
                bool good = true;
                // only bytes of this group are read (the next group may still be inflating): a block_size field cut
                // by the group's end is left to the serial pass
                while (p + 4 <= x.end) {
                    const int32_t bs = rdi32(d + p);
                    if (bs < 32) { good = false; break; }
                    if (p + 4 + (size_t)bs > ulen) { x.wave_end = true; break; }   // the wave's last, partial record
                    x.recs.push_back({p + 4, (uint32_t)bs});
                    p += 4 + (size_t)bs;
                }
                x.exit = p;
                x.ok = good;
            };
            std::atomic<size_t> next(0);
            std::atomic<bool> ok(true);
            auto work = [&]() {
                for (;;) {
                    const size_t g = next.fetch_add(1);
                    if (g >= n_groups) break;
                    bool fine = true;
                    for (size_t b = g * gb; b < std::min(blocks.size(), (g + 1) * gb); ++b)
                        if (!inflate_block(cbase + blocks[b].coff, blocks[b].csize, ubuf.get() + base + blocks[b].uoff, blocks[b].usize))
                            fine = false;
                    if (!fine) ok = false;
                    else if (n_groups > 1) walk(st[g]);
                }
            };
            const int nt = (int)std::min<size_t>((size_t)n_threads, std::max<size_t>(1, n_groups));
            std::vector<std::thread> th;
            for (int t = 1; t < nt; ++t) th.emplace_back(work);
            work();
            for (auto &t : th) t.join();
            if (!ok) { hx_set_error("inflate failed on %s", path.c_str()); return HX_E_ARG; }
        }
        auto t2 = clk::now();
        size_t q = 0;
        if (!header_done) {
            const long used = parse_header(w);
            if (used < 0) return HX_E_ARG;
            if (used == 0) {                     // header longer than this wave: keep everything, read on
                if (eof) { hx_set_error("%s: truncated BAM header", path.c_str()); return HX_E_ARG; }
                carry.assign(ubuf.get(), ubuf.get() + ulen);
                return HX_OK;
            }
            q = (size_t)used;
        }
        const uint8_t *d = ubuf.get();
        size_t si = 0;
        bool at_end = false;
        while (!at_end && q + 4 <= ulen) {
            while (si < st.size() && st[si].start < q) ++si;       // stretches the chain has already passed
            if (si < st.size() && st[si].start == q) {
                if (st[si].ok) {
                    recs.insert(recs.end(), st[si].recs.begin(), st[si].recs.end());
                    at_end = st[si].wave_end || st[si].exit + 4 > ulen;            // stopped at the wave's partial record
                    q = st[si].exit;
                    ++si;
                    continue;
                }
                ++si;                                  // it met a block_size < 32: the walk below says what is wrong
            }
            const size_t stop = si < st.size() ? st[si].start : ulen;     // walk up to the next stretch
            while (q < stop && q + 4 <= ulen) {
                const int32_t bs = rdi32(d + q);
                if (bs < 32) { hx_set_error("%s: corrupt alignment record (block_size %d)", path.c_str(), bs); return HX_E_ARG; }
                if (q + 4 + (size_t)bs > ulen) { at_end = true; break; }
                __builtin_prefetch(d + std::min(ulen - 1, q + 16 * (4 + (size_t)bs)));
                recs.push_back({q + 4, (uint32_t)bs});
                q += 4 + (size_t)bs;
            }
            if (q >= ulen || q + 4 > ulen) break;
        }
        carry.assign(d + q, d + ulen);
        if (eof && !carry.empty()) { hx_set_error("%s: truncated BAM (partial record at EOF)", path.c_str()); return HX_E_ARG; }
        auto t3 = clk::now();
        t_read += std::chrono::duration<double>(t1 - t0).count();
        t_inflate += std::chrono::duration<double>(t2 - t1).count();
        t_scan += std::chrono::duration<double>(t3 - t2).count();
        return HX_OK;
    }

    // a coordinate-sorted file has nothing more for [.., end_pos] on the target once its records are beyond it
    bool past_region(int32_t end_pos) const { return past_region(own, end_pos); }
    bool past_region(const Wave &w, int32_t end_pos) const {
        if (!sorted || w.recs.empty()) return false;
        const uint8_t *r = w.ubuf.get() + w.recs.back().off;
        const int32_t tid = rdi32(r), pos = rdi32(r + 4);
        if (tid < 0) return true;                                  // unmapped reads without a position come last
        return tid > target_tid || (tid == target_tid && pos + 1 > end_pos);
    }
};

constexpr size_t WAVE_CBYTES = (size_t)48 << 20;      // compressed bytes per wave (~150-250 MB inflated)

}  // namespace

extern "C" {

int hx_pack_bam(const char *bam_path, const char *contig, int32_t start_pos, int32_t end_pos,
                const int32_t *snp_pos, int32_t n_snps, int stepper, int n_threads, hx_packed *out) {
    return hx_pack_bam_ex(bam_path, contig, start_pos, end_pos, snp_pos, n_snps, stepper, n_threads, 0, out, nullptr);
}

int hx_pack_bam_ex(const char *bam_path, const char *contig, int32_t start_pos, int32_t end_pos,
                   const int32_t *snp_pos, int32_t n_snps, int stepper, int n_threads, int32_t max_depth,
                   hx_packed *out, double stage_seconds[6]) {
    if (!bam_path || !contig || !out || n_snps < 0 || (n_snps && !snp_pos) || stepper < 0 || stepper > 2 || max_depth < 0) {
        hx_set_error("hx_pack_bam: bad arguments");
        return HX_E_ARG;
    }
    for (int32_t i = 1; i < n_snps; ++i)
        if (snp_pos[i] <= snp_pos[i - 1]) { hx_set_error("hx_pack_bam: SNP positions must be strictly increasing"); return HX_E_ARG; }
    memset(out, 0, sizeof(*out));
    if (n_threads < 1) n_threads = 1;
    using clk = std::chrono::steady_clock;
    BamStream bs;
    int rc = bs.open(bam_path, contig, n_threads);
    if (rc) return rc;
    std::vector<std::vector<Out>> waves;           // per wave, per thread range (BAM order is kept)
    double t_walk = 0, t_depth = 0, t_gather = 0;
    int64_t n_records = 0;
    // pysam's pileup engine (bam_plp) buffers at most max_depth reads: a read that is not the first of its start
    // position is dropped while max_depth reads are still live (gretel never changes pysam's default of 8000).
    // Reads arrive sorted by start, so "live" = admitted - retired with the ends counted in a ring indexed by position.
    std::vector<uint32_t> end_ring;
    int64_t ring_mask = 0, depth_pos = -1, retired_upto = -1, live = 0;
    std::vector<int64_t> rbeg, rend;
    std::vector<uint8_t> admit;
    // two waves take turns: while this thread filters and walks wave i, a helper reads, inflates and indexes wave i+1
    // (the serial parts of either side - the record scan, the depth cap - hide behind the other side's parallel work)
    Wave wv[2];
    // compressed bytes per wave: small enough that the first wave's inflate and the last wave's walk - the two ends of
    // the pipeline that nothing hides - stay short (HX_PACK_WAVE_MB overrides, for measurements)
    static const size_t PACK_WAVE = [] {
        const char *e = getenv("HX_PACK_WAVE_MB");
        const long mb = e ? atol(e) : 16;
        return (size_t)(mb < 1 ? 1 : mb > 256 ? 256 : mb) << 20;
    }();
    std::future<int> next = std::async(std::launch::async, [&]() { return bs.next_wave(wv[0], PACK_WAVE); });
    for (int wi = 0;; ++wi) {
        rc = next.get();
        if (rc) return rc;
        Wave &cw = wv[wi & 1];
        const bool more = !(bs.eof || (bs.header_done && bs.past_region(cw, end_pos)));
        if (more) next = std::async(std::launch::async, [&, wi]() { return bs.next_wave(wv[(wi + 1) & 1], PACK_WAVE); });
        const size_t nrec = cw.recs.size();
        n_records += (int64_t)nrec;
        const uint8_t *d = cw.ubuf.get();
        const std::vector<RecRef> &recs = cw.recs;
        if (nrec) {
            auto t0 = clk::now();
            const bool depth_on = max_depth > 0;
            if (depth_on) {
                admit.assign(nrec, 1);
                rbeg.resize(nrec); rend.resize(nrec);
                // only reads the pileup iterator fetches and its stepper lets through count: target contig,
                // overlapping [start_pos - 1, end_pos); their spans are extracted in parallel (end = -1: not counted)
                {
                    const int ntd = (int)std::min<size_t>((size_t)n_threads, std::max<size_t>(1, nrec / 4096));
                    std::atomic<int64_t> span_max(0);
                    auto spans = [&](int t) {
                        int64_t mx = 0;
                        const size_t a = nrec * (size_t)t / (size_t)ntd, b = nrec * (size_t)(t + 1) / (size_t)ntd;
                        for (size_t i = a; i < b; ++i) {
                            const Rec r = parse_rec(d + recs[i].off, recs[i].size);
                            rend[i] = -1;
                            if (!r.ok || r.tid != bs.target_tid || r.pos < 0 || !passes_stepper(r.flag, stepper)) continue;
                            const int64_t beg = r.pos, end = (int64_t)r.pos + std::max<int64_t>(1, ref_len_of(r));
                            if (beg >= end_pos || end <= (int64_t)start_pos - 1) continue;
                            rbeg[i] = beg; rend[i] = end;
                            mx = std::max(mx, end - beg);
                        }
                        int64_t cur = span_max.load();
                        while (mx > cur && !span_max.compare_exchange_weak(cur, mx)) {}
                    };
                    std::vector<std::thread> th;
                    for (int t = 1; t < ntd; ++t) th.emplace_back(spans, t);
                    spans(0);
                    for (auto &t : th) t.join();
                    // the ring must hold every end that is still ahead of the sweep
                    int64_t need = 2;
                    while (need < span_max.load() + 2) need <<= 1;
                    if (need > (int64_t)end_ring.size()) {
                        std::vector<uint32_t> bigger((size_t)need, 0u);
                        for (int64_t pos = retired_upto + 1; !end_ring.empty() && pos <= retired_upto + (int64_t)end_ring.size(); ++pos)
                            bigger[(size_t)(pos & (need - 1))] = end_ring[(size_t)(pos & ring_mask)];
                        end_ring.swap(bigger);
                        ring_mask = need - 1;
                    }
                }
                for (size_t i = 0; i < nrec; ++i) {
                    if (rend[i] < 0) continue;
                    const int64_t beg = rbeg[i];
                    if (beg != depth_pos) {
                        // the columns before the new start position have been emitted: reads ending there are gone
                        const int64_t upto = beg - 1;
                        if (live == 0) retired_upto = std::max(retired_upto, upto);
                        for (; retired_upto < upto; ) {
                            ++retired_upto;
                            uint32_t &c = end_ring[(size_t)(retired_upto & ring_mask)];
                            live -= c; c = 0;
                        }
                        depth_pos = beg;
                    } else if (live >= max_depth) {
                        admit[i] = 0;
                        continue;
                    }
                    end_ring[(size_t)(rend[i] & ring_mask)]++;
                    ++live;
                }
            }
            auto t1 = clk::now();
            const int nt = (int)std::min<size_t>((size_t)n_threads, std::max<size_t>(1, nrec / 4096));
            waves.emplace_back((size_t)nt);
            std::vector<Out> &outs = waves.back();
            auto work = [&](int t) {
                std::vector<uint8_t> tmp;
                Out mine;                             // (the Out objects of neighbouring threads share cache lines, and
                                                      //  every push_back writes the vector's end pointer)
                const size_t a = nrec * (size_t)t / (size_t)nt, b = nrec * (size_t)(t + 1) / (size_t)nt;
                mine.rank.reserve((b - a) / 2 + 16);
                mine.klen.reserve((b - a) / 2 + 16);
                mine.codes.reserve((b - a) * 8 + 64);
                for (size_t i = a; i < b; ++i) {
                    if (depth_on && !admit[i]) continue;
                    pack_record(parse_rec(d + recs[i].off, recs[i].size), bs.target_tid, start_pos, end_pos, snp_pos,
                                n_snps, stepper, mine, tmp);
                }
                outs[(size_t)t] = std::move(mine);
            };
            std::vector<std::thread> th;
            for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
            work(0);
            for (auto &t : th) t.join();
            auto t2 = clk::now();
            t_depth += std::chrono::duration<double>(t1 - t0).count();
            t_walk += std::chrono::duration<double>(t2 - t1).count();
        }
        if (!more) break;
    }
    auto t0 = clk::now();
    // gather: prefix the per-range sizes, then every range is copied into place by its own thread
    std::vector<Out *> parts;
    for (auto &w : waves) for (auto &o : w) parts.push_back(&o);
    std::vector<int64_t> r0(parts.size() + 1, 0), c0(parts.size() + 1, 0);
    for (size_t i = 0; i < parts.size(); ++i) {
        r0[i + 1] = r0[i] + (int64_t)parts[i]->rank.size();
        c0[i + 1] = c0[i] + (int64_t)parts[i]->codes.size();
    }
    const int64_t R = r0.back(), C = c0.back();
    out->rank = (int32_t *)malloc(sizeof(int32_t) * (size_t)std::max<int64_t>(R, 1));
    out->off = (int64_t *)malloc(sizeof(int64_t) * (size_t)(R + 1));
    out->codes = (uint8_t *)malloc((size_t)std::max<int64_t>(C, 1) + 16);
    if (!out->rank || !out->off || !out->codes) { hx_pack_free(out); return HX_E_NOMEM; }
    {
        std::atomic<size_t> next(0);
        auto work = [&]() {
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= parts.size()) break;
                const Out &o = *parts[i];
                if (!o.rank.empty()) memcpy(out->rank + r0[i], o.rank.data(), sizeof(int32_t) * o.rank.size());
                int64_t acc = c0[i];
                for (size_t j = 0; j < o.klen.size(); ++j) { out->off[r0[i] + (int64_t)j] = acc; acc += o.klen[j]; }
                if (!o.codes.empty()) memcpy(out->codes + c0[i], o.codes.data(), o.codes.size());
            }
        };
        const int nt = (int)std::min<size_t>((size_t)n_threads, std::max<size_t>(1, parts.size()));
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(work);
        work();
        for (auto &t : th) t.join();
    }
    out->off[R] = C;
    t_gather = std::chrono::duration<double>(clk::now() - t0).count();
    out->n_reads = R;
    out->n_codes = C;
    out->n_records = n_records;
    if (stage_seconds) {
        stage_seconds[0] = bs.t_read; stage_seconds[1] = bs.t_inflate; stage_seconds[2] = bs.t_scan;
        stage_seconds[3] = t_depth; stage_seconds[4] = t_walk; stage_seconds[5] = t_gather;
    }
    if (getenv("HX_PACK_VERBOSE"))
        fprintf(stderr, "[hx_pack_bam] read %.1f ms, inflate %.1f ms, scan %.1f ms, depth %.1f ms, walk %.1f ms, gather %.1f ms\n",
                1e3 * bs.t_read, 1e3 * bs.t_inflate, 1e3 * bs.t_scan, 1e3 * t_depth, 1e3 * t_walk, 1e3 * t_gather);
    return HX_OK;
}

/* Per-position A,C,G,T counts over [start0, end0) of a contig from every alignment (no filter, no base
 * quality threshold): what gretel/snpper.py:30 asks pysam's count_coverage for.  out[4][end0-start0]. */
int hx_count_coverage(const char *bam_path, const char *contig, int32_t start0, int32_t end0, int n_threads,
                      uint32_t *out) {
    if (!bam_path || !contig || !out || start0 < 0 || end0 < start0) { hx_set_error("hx_count_coverage: bad arguments"); return HX_E_ARG; }
    if (n_threads < 1) n_threads = 1;
    BamStream bs;
    int rc = bs.open(bam_path, contig, n_threads);
    if (rc) return rc;
    const int64_t len = (int64_t)end0 - start0;
    memset(out, 0, sizeof(uint32_t) * 4 * (size_t)len);
    std::vector<std::vector<uint32_t>> part((size_t)n_threads);
    for (;;) {
        rc = bs.next_wave(WAVE_CBYTES);
        if (rc) return rc;
        const size_t nrec = bs.recs.size();
        const uint8_t *d = bs.ubuf.get();
        if (nrec) {
            const int nt = (int)std::min<size_t>((size_t)n_threads, std::max<size_t>(1, nrec / 4096));
            auto work = [&](int t) {
                std::vector<uint32_t> &c = part[(size_t)t];
                if (c.empty()) c.assign(4 * (size_t)len, 0u);
                const size_t a = nrec * (size_t)t / (size_t)nt, b = nrec * (size_t)(t + 1) / (size_t)nt;
                for (size_t i = a; i < b; ++i) {
                    const Rec r = parse_rec(d + bs.recs[i].off, bs.recs[i].size);
                    if (!r.ok || r.tid != bs.target_tid || r.pos < 0) continue;
                    int64_t rpos = r.pos, qpos = 0;
                    for (int ci = 0; ci < r.n_cigar; ++ci) {
                        const uint32_t cv = rd32(r.cig + 4 * (size_t)ci);
                        const int op = cv & 0xf;
                        const int64_t ln = cv >> 4;
                        if (op == 0 || op == 7 || op == 8) {
                            for (int64_t j = 0; j < ln; ++j) {
                                const int64_t rp = rpos + j, q = qpos + j;
                                if (rp < start0 || rp >= end0 || q >= r.l_seq) continue;
                                const uint8_t bb = r.seq[q >> 1];
                                const uint8_t code = NT16_CODE[(q & 1) ? (bb & 0xf) : (bb >> 4)];
                                if (code < 4) c[(size_t)code * (size_t)len + (size_t)(rp - start0)]++;
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
        }
        if (bs.eof || (bs.header_done && bs.past_region(end0))) break;
    }
    for (auto &c : part)
        if (!c.empty())
            for (size_t i = 0; i < 4 * (size_t)len; ++i) out[i] += c[i];
    return HX_OK;
}

/* The same counts with the histogram on the GPU (coverage.cu): the CPU decodes the BAM wave by wave into aligned
 * segments + the BAM's own 4-bit bases, the device counts them.  out[4*(end0-start0)] on the host. */
int hx_count_coverage_gpu(const char *bam_path, const char *contig, int32_t start0, int32_t end0, int n_threads,
                          int32_t device, uint32_t *out) {
    if (!bam_path || !contig || !out || start0 < 0 || end0 < start0) { hx_set_error("hx_count_coverage_gpu: bad arguments"); return HX_E_ARG; }
    if (n_threads < 1) n_threads = 1;
    const int64_t len = (int64_t)end0 - start0;
    if (len == 0) return HX_OK;
    BamStream bs;
    int rc = bs.open(bam_path, contig, n_threads);
    if (rc) return rc;
#define COV_CUDA(call)                                                                                    \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess) {                                                                         \
            hx_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));           \
            for (void *p__ : dev_bufs) if (p__) cudaFree(p__);                                            \
            if (st) cudaStreamDestroy(st);                                                                \
            return HX_E_CUDA;                                                                             \
        }                                                                                                 \
    } while (0)
    cudaStream_t st = nullptr;
    uint32_t *d_counts = nullptr;
    int32_t *d_start = nullptr, *d_len = nullptr;
    int64_t *d_nib = nullptr;
    uint8_t *d_seq = nullptr;
    size_t cap_seg = 0, cap_seq = 0;
    std::vector<void *> dev_bufs;
    COV_CUDA(cudaSetDevice(device));
    COV_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    COV_CUDA(cudaMalloc((void **)&d_counts, sizeof(uint32_t) * 4 * (size_t)len));
    dev_bufs.push_back(d_counts);
    COV_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(uint32_t) * 4 * (size_t)len, st));
    struct Part { std::vector<int32_t> start, len; std::vector<int64_t> nib; std::vector<uint8_t> seq; };
    std::vector<int32_t> h_start, h_len;
    std::vector<int64_t> h_nib;
    std::vector<uint8_t> h_seq;
    for (;;) {
        rc = bs.next_wave(WAVE_CBYTES / 4);
        if (rc) break;
        const size_t nrec = bs.recs.size();
        const uint8_t *d = bs.ubuf.get();
        if (nrec) {
            const int nt = (int)std::min<size_t>((size_t)n_threads, std::max<size_t>(1, nrec / 4096));
            std::vector<Part> parts((size_t)nt);
            auto work = [&](int t) {
                Part &P = parts[(size_t)t];
                const size_t a = nrec * (size_t)t / (size_t)nt, b = nrec * (size_t)(t + 1) / (size_t)nt;
                for (size_t i = a; i < b; ++i) {
                    const Rec r = parse_rec(d + bs.recs[i].off, bs.recs[i].size);
                    if (!r.ok || r.tid != bs.target_tid || r.pos < 0 || r.l_seq <= 0) continue;
                    const int64_t seq_nib0 = 2 * (int64_t)P.seq.size();
                    bool used = false;
                    int64_t rpos = r.pos, qpos = 0;
                    for (int ci = 0; ci < r.n_cigar; ++ci) {
                        const uint32_t cv = rd32(r.cig + 4 * (size_t)ci);
                        const int op = cv & 0xf;
                        const int64_t ln = cv >> 4;
                        if (op == 0 || op == 7 || op == 8) {
                            const int64_t usable = std::min<int64_t>(ln, (int64_t)r.l_seq - qpos);
                            if (usable > 0 && rpos < end0 && rpos + usable > start0) {
                                P.start.push_back((int32_t)rpos);
                                P.len.push_back((int32_t)usable);
                                P.nib.push_back(seq_nib0 + qpos);
                                used = true;
                            }
                            rpos += ln; qpos += ln;
                        } else if (op == 2 || op == 3) {
                            rpos += ln;
                        } else if (op == 1 || op == 4) {
                            qpos += ln;
                        }
                    }
                    if (used) P.seq.insert(P.seq.end(), r.seq, r.seq + ((size_t)r.l_seq + 1) / 2);
                    else { /* nothing shipped for this record */ }
                }
            };
            std::vector<std::thread> th;
            for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
            work(0);
            for (auto &t : th) t.join();
            // gather (the nibble offsets of a part are relative to its own bases)
            h_start.clear(); h_len.clear(); h_nib.clear(); h_seq.clear();
            for (auto &P : parts) {
                const int64_t shift = 2 * (int64_t)h_seq.size();
                h_start.insert(h_start.end(), P.start.begin(), P.start.end());
                h_len.insert(h_len.end(), P.len.begin(), P.len.end());
                for (int64_t v : P.nib) h_nib.push_back(v + shift);
                h_seq.insert(h_seq.end(), P.seq.begin(), P.seq.end());
            }
            const size_t ns = h_start.size();
            if (ns) {
                COV_CUDA(cudaStreamSynchronize(st));               // the previous wave's kernel has read its buffers
                if (ns > cap_seg) {
                    for (void *p : {(void *)d_start, (void *)d_len, (void *)d_nib}) if (p) cudaFree(p);
                    dev_bufs.resize(1);
                    if (d_seq) dev_bufs.push_back(d_seq);
                    cap_seg = ns + ns / 4;
                    d_start = nullptr; d_len = nullptr; d_nib = nullptr;
                    COV_CUDA(cudaMalloc((void **)&d_start, 4 * cap_seg)); dev_bufs.push_back(d_start);
                    COV_CUDA(cudaMalloc((void **)&d_len, 4 * cap_seg)); dev_bufs.push_back(d_len);
                    COV_CUDA(cudaMalloc((void **)&d_nib, 8 * cap_seg)); dev_bufs.push_back(d_nib);
                }
                if (h_seq.size() + 16 > cap_seq) {
                    if (d_seq) { cudaFree(d_seq); dev_bufs.erase(std::find(dev_bufs.begin(), dev_bufs.end(), (void *)d_seq)); }
                    cap_seq = h_seq.size() + h_seq.size() / 4 + 16;
                    d_seq = nullptr;
                    COV_CUDA(cudaMalloc((void **)&d_seq, cap_seq)); dev_bufs.push_back(d_seq);
                }
                COV_CUDA(cudaMemcpyAsync(d_start, h_start.data(), 4 * ns, cudaMemcpyHostToDevice, st));
                COV_CUDA(cudaMemcpyAsync(d_len, h_len.data(), 4 * ns, cudaMemcpyHostToDevice, st));
                COV_CUDA(cudaMemcpyAsync(d_nib, h_nib.data(), 8 * ns, cudaMemcpyHostToDevice, st));
                COV_CUDA(cudaMemcpyAsync(d_seq, h_seq.data(), h_seq.size(), cudaMemcpyHostToDevice, st));
                rc = hx_launch_coverage(d_start, d_nib, d_len, d_seq, (int64_t)ns, start0, (int32_t)len, d_counts, st);
                if (rc) break;
                COV_CUDA(cudaStreamSynchronize(st));               // the host vectors are reused by the next wave
            }
        }
        if (bs.eof || (bs.header_done && bs.past_region(end0))) break;
    }
    if (!rc) {
        COV_CUDA(cudaMemcpyAsync(out, d_counts, sizeof(uint32_t) * 4 * (size_t)len, cudaMemcpyDeviceToHost, st));
        COV_CUDA(cudaStreamSynchronize(st));
    }
    for (void *p : dev_bufs) if (p) cudaFree(p);
    cudaStreamDestroy(st);
#undef COV_CUDA
    return rc;
}

int hx_inflate_raw(const uint8_t *src, int64_t n, int64_t n_readable, uint8_t *dst, int64_t m, int32_t use_zlib) {
    if (!src || !dst || n < 0 || m < 0 || n_readable < n) { hx_set_error("hx_inflate_raw: bad arguments"); return HX_E_ARG; }
    const bool ok = use_zlib ? (m == 0 || inflate_block_zlib(src, (size_t)n, dst, (size_t)m))
                             : hxz::inflate_raw(src, (size_t)n, (size_t)n_readable, dst, (size_t)m);
    if (!ok) { hx_set_error("hx_inflate_raw: not a DEFLATE stream of %lld bytes inflating to %lld", (long long)n, (long long)m); return HX_E_ARG; }
    return HX_OK;
}

int hx_bam_contig_length(const char *bam_path, const char *contig, int32_t *length) {
    if (!bam_path || !contig || !length) { hx_set_error("hx_bam_contig_length: bad arguments"); return HX_E_ARG; }
    BamStream bs;
    int rc = bs.open(bam_path, contig, 1);
    if (rc) return rc;
    while (!bs.header_done) {                      // only the header is read: 1 MiB at a time
        rc = bs.next_wave((size_t)1 << 20);
        if (rc) return rc;
        if (bs.eof && !bs.header_done) { hx_set_error("%s: truncated BAM header", bam_path); return HX_E_ARG; }
    }
    *length = bs.target_len;
    return HX_OK;
}

// ---- dense wire format encoder (CPU side of hx_ingest_host_dense; layout documented in wire.cu) ----------
}  // extern "C"

#include "dense_enc.h"

static inline int64_t al16(int64_t x) { return (x + 15) & ~(int64_t)15; }

static inline uint64_t ld64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

// A small persistent pool: the per-chunk encoding passes last a few hundred microseconds, which is what creating
// sixteen threads costs.  One job at a time (callers serialise on the mutex); the workers are detached and sleep on
// a condition variable between jobs.
#include <condition_variable>
#include <functional>
#include <mutex>
namespace {
struct HxPool {
    std::mutex call_mu, mu;
    std::condition_variable cv_go, cv_done;
    std::function<void(int)> job;
    int n_workers = 0, want = 0, pending = 0;
    uint64_t gen = 0;
    void worker(int id) {
        uint64_t seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu);
            cv_go.wait(lk, [&] { return gen != seen; });
            seen = gen;
            if (id >= want) continue;
            std::function<void(int)> f = job;
            lk.unlock();
            f(id);
            lk.lock();
            if (--pending == 0) cv_done.notify_one();
        }
    }
    void run(int nt, const std::function<void(int)> &f) {
        std::lock_guard<std::mutex> call(call_mu);
        if (nt <= 1) { f(0); return; }
        {
            std::unique_lock<std::mutex> lk(mu);
            while (n_workers < nt - 1) {
                const int id = ++n_workers;                       // worker ids 1..: id 0 is the caller
                std::thread([this, id] { worker(id); }).detach();
            }
            job = f;
            want = nt;
            pending = nt - 1;
            ++gen;
        }
        cv_go.notify_all();
        f(0);
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return pending == 0; });
    }
};
HxPool *hx_pool() { static HxPool *p = new HxPool(); return p; }   // leaked on purpose: workers outlive static destructors
}  // namespace

template <class F>
static void run_threads(int nt, F f) {
    hx_pool()->run(nt, std::function<void(int)>(f));
}

#if defined(__x86_64__)
#include <immintrin.h>
// 0 = 64-bit words, 1 = SSE4.1 (16 alleles per step), 2 = AVX2 (32 per step): what the CPU supports, capped by
// HX_DENSE_SIMD=0|1|2 (the tests run every level the machine has against each other)
static int dense_simd_level() {
    int have = __builtin_cpu_supports("avx2") ? 2 : (__builtin_cpu_supports("sse4.1") && __builtin_cpu_supports("ssse3")) ? 1 : 0;
    if (const char *e = getenv("HX_DENSE_SIMD")) have = std::min(have, std::max(0, atoi(e)));
    return have;
}
// alleles >= 4 (N, -, _) among the bytes flagged in m are listed as exceptions (inside [xlo, xhi) only: the caller's own reads)
static inline void dense_note_mask(uint32_t m, int64_t i, int64_t xlo, int64_t xhi, std::vector<uint32_t> &exc) {
    while (m) {
        const int64_t j = i + __builtin_ctz(m);
        if (j >= xlo && j < xhi) exc.push_back((uint32_t)j);
        m &= m - 1;
    }
}
// 32 alleles -> 8 bytes per step (allele j of a byte at bits 2j, as the 64-bit loop of hx_dense_pack); a code > 6 sets
// `any`.  Returns where it stopped.
__attribute__((target("avx2"))) static int64_t dense_pack_avx2(const uint8_t *codes, uint8_t *c2, int64_t i, int64_t hi,
                                                                int64_t xlo, int64_t xhi, std::vector<uint32_t> &exc,
                                                                uint64_t &any) {
    const __m256i three = _mm256_set1_epi8(3), six = _mm256_set1_epi8(6);
    const __m256i w14 = _mm256_set1_epi16(0x0401), w116 = _mm256_set1_epi32(0x00100001);
    for (; i + 32 <= hi; i += 32) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(codes + i));
        const uint32_t sign = (uint32_t)_mm256_movemask_epi8(v);                              // codes >= 128
        const uint32_t m = (uint32_t)_mm256_movemask_epi8(_mm256_cmpgt_epi8(v, three)) | sign;      // codes >= 4
        if (m) {
            if (sign | (uint32_t)_mm256_movemask_epi8(_mm256_cmpgt_epi8(v, six))) any = 1;
            dense_note_mask(m, i, xlo, xhi, exc);
        }
        const __m256i a = _mm256_and_si256(v, three);
        const __m256i t = _mm256_madd_epi16(_mm256_maddubs_epi16(a, w14), w116);     // a byte of four alleles per 32-bit lane
        const __m256i p16 = _mm256_packus_epi32(t, t);
        const __m256i p8 = _mm256_packus_epi16(p16, p16);
        const uint32_t lo4 = (uint32_t)_mm256_extract_epi32(p8, 0), hi4 = (uint32_t)_mm256_extract_epi32(p8, 4);
        const uint64_t out = (uint64_t)lo4 | ((uint64_t)hi4 << 32);
        memcpy(c2 + (i >> 2), &out, 8);
    }
    return i;
}
// the same, 2 x 16 alleles per step
__attribute__((target("sse4.1,ssse3"))) static int64_t dense_pack_sse41(const uint8_t *codes, uint8_t *c2, int64_t i, int64_t hi,
                                                                        int64_t xlo, int64_t xhi, std::vector<uint32_t> &exc,
                                                                        uint64_t &any) {
    const __m128i three = _mm_set1_epi8(3), six = _mm_set1_epi8(6);
    const __m128i w14 = _mm_set1_epi16(0x0401), w116 = _mm_set1_epi32(0x00100001);
    for (; i + 32 <= hi; i += 32) {
        const __m128i v0 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(codes + i));
        const __m128i v1 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(codes + i + 16));
        const uint32_t sign = (uint32_t)_mm_movemask_epi8(v0) | ((uint32_t)_mm_movemask_epi8(v1) << 16);
        const uint32_t m = (uint32_t)_mm_movemask_epi8(_mm_cmpgt_epi8(v0, three)) |
                           ((uint32_t)_mm_movemask_epi8(_mm_cmpgt_epi8(v1, three)) << 16) | sign;
        if (m) {
            if (sign | (uint32_t)_mm_movemask_epi8(_mm_or_si128(_mm_cmpgt_epi8(v0, six), _mm_cmpgt_epi8(v1, six)))) any = 1;
            dense_note_mask(m, i, xlo, xhi, exc);
        }
        const __m128i t0 = _mm_madd_epi16(_mm_maddubs_epi16(_mm_and_si128(v0, three), w14), w116);
        const __m128i t1 = _mm_madd_epi16(_mm_maddubs_epi16(_mm_and_si128(v1, three), w14), w116);
        const __m128i p16 = _mm_packus_epi32(t0, t1);
        const __m128i p8 = _mm_packus_epi16(p16, p16);
        _mm_storel_epi64(reinterpret_cast<__m128i *>(c2 + (i >> 2)), p8);
    }
    return i;
}
#endif

int hx_dense_begin(const int64_t *off, int64_t n_reads, int n_threads, int64_t kmax_hint, HxDensePlan *pl, bool slim) {
    HxDensePlan &P = *pl;
    P.slim = slim;
    P.c0 = n_reads ? off[0] : 0;
    P.n_codes = n_reads ? off[n_reads] - P.c0 : 0;
    P.n_reads = n_reads;
    if (P.n_codes < 0 || P.n_codes >= ((int64_t)1 << 32)) {
        hx_set_error("hx_dense_encode: %lld alleles in one chunk (limit 2^32 - 1): split the reads", (long long)P.n_codes);
        return HX_E_ARG;
    }
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(n_threads, HX_DENSE_MAX_THREADS), 1 + n_reads / 65536));
    P.nt = nt;
    int64_t km = kmax_hint;
    if (km <= 0) {
        std::vector<int64_t> kmax((size_t)nt, 0);
        run_threads(nt, [&](int t) {
            int64_t m = 0;
            for (int64_t r = n_reads * t / nt; r < n_reads * (t + 1) / nt; ++r) m = std::max(m, off[r + 1] - off[r]);
            kmax[(size_t)t] = m;
        });
        for (int64_t m : kmax) km = std::max(km, m);
    }
    if (km > 65535) { hx_set_error("hx_dense_encode: a read covers more than 65535 SNPs"); return HX_E_ARG; }
    P.klen_bytes = km < 256 ? 1 : 2;
    const int64_t n_words = slim ? 0 : (P.n_codes + 15) / 16;
    P.o_klen = al16(n_reads);
    P.o_codes2 = P.o_klen + al16(n_reads * P.klen_bytes);
    P.o_exc = P.o_codes2 + al16(n_words * 4);
    P.head_bytes = P.o_exc;
    P.bytes = 0;
    return HX_OK;
}

// The one pass over the reads.  Eight alleles per 64-bit operation; every byte of the fixed sections that the device
// reads is written (the blob need not be zeroed).
int hx_dense_pack(const int32_t *rank, const int64_t *off, const uint8_t *codes, HxDensePlan *pl, uint8_t *blob) {
    HxDensePlan &P = *pl;
    const int nt = P.nt;
    const int64_t n_reads = P.n_reads, n_codes = P.n_codes, c0 = P.c0;
    const int kb = P.klen_bytes;
    const int64_t klim = kb == 1 ? 255 : 65535;
    uint8_t *c2 = blob + P.o_codes2;
    std::vector<int> bad((size_t)nt, 0);
#if defined(__x86_64__)
    const int simd = P.slim ? 0 : dense_simd_level();
#endif
    run_threads(nt, [&](int t) {
        const int64_t a = n_reads * t / nt, b = n_reads * (t + 1) / nt;
        std::vector<uint32_t> &exc = P.exc[t];
        std::vector<int64_t> &ei = P.esc_idx[t];
        std::vector<int32_t> &ed = P.esc_delta[t];
        exc.clear(); ei.clear(); ed.clear();
        int bd = 0;
        {
            // ranks and SNP counts: a branch-free (vectorisable) pass; the rare escapes (a gap of >= 255 sites, the
            // chunk's first read) are collected in a second look at the blocks that hold one
            constexpr int64_t BLK = 4096;
            for (int64_t r0 = a; r0 < b; r0 += BLK) {
                const int64_t r1 = std::min(b, r0 + BLK);
                int64_t neg = 0, kbad = 0, esc = 0;
                int64_t r = r0;
                if (r == 0) {
                    const int64_t d = rank[0], k = off[1] - off[0];
                    neg |= d; kbad |= (k < 0) | (k > klim); esc |= d >= 255;
                    blob[0] = (uint8_t)std::min<int64_t>(std::max<int64_t>(d, 0), 255);
                    if (kb == 1) blob[P.o_klen] = (uint8_t)k; else ((uint16_t *)(blob + P.o_klen))[0] = (uint16_t)k;
                    r = 1;
                }
                if (kb == 1) {
                    uint8_t *kl = blob + P.o_klen;
                    for (; r < r1; ++r) {
                        const int32_t d = rank[r] - rank[r - 1];
                        const int64_t k = off[r + 1] - off[r];
                        neg |= d; kbad |= (k < 0) | (k > 255); esc |= d >= 255;
                        blob[r] = (uint8_t)(d > 255 ? 255 : d);
                        kl[r] = (uint8_t)k;
                    }
                } else {
                    uint16_t *kl = (uint16_t *)(blob + P.o_klen);
                    for (; r < r1; ++r) {
                        const int32_t d = rank[r] - rank[r - 1];
                        const int64_t k = off[r + 1] - off[r];
                        neg |= d; kbad |= (k < 0) | (k > 65535); esc |= d >= 255;
                        blob[r] = (uint8_t)(d > 255 ? 255 : d);
                        kl[r] = (uint16_t)k;
                    }
                }
                if (neg < 0) bd = 1;
                if (kbad) bd = 2;
                if (esc)
                    for (int64_t q = r0; q < r1; ++q) {
                        const int64_t d = (int64_t)rank[q] - (q ? (int64_t)rank[q - 1] : 0);
                        if (d >= 255) { ei.push_back(q); ed.push_back((int32_t)d); }
                    }
            }
        }
        if (P.slim || b <= a || bd) { bad[(size_t)t] = bd; return; }
        // This thread lists the exceptions of its reads' alleles [xlo, xhi) and writes the output bytes of the
        // alleles [lo, hi): the same range rounded so that every output byte has exactly one writer.
        const int64_t xlo = off[a] - c0, xhi = off[b] - c0;
        int64_t lo = t == 0 ? 0 : (xlo + 3) & ~(int64_t)3;
        int64_t hi = t == nt - 1 ? n_codes : (xhi + 3) & ~(int64_t)3;
        hi = std::min(hi, n_codes);
        uint64_t any = 0;
        auto note = [&](int64_t i, uint64_t x) {       // exceptions among the eight alleles at i, inside [xlo, xhi)
            uint64_t m = x & 0x0404040404040404ull;
            while (m) {
                const int64_t j = i + (__builtin_ctzll(m) >> 3);
                if (j >= xlo && j < xhi) exc.push_back((uint32_t)j);
                m &= m - 1;
            }
        };
        // head: the alleles of this thread's reads before its first output byte belong to the previous writer
        for (int64_t i = xlo; i < std::min(lo, xhi); ++i) { const uint8_t c = codes[c0 + i]; if (c > 6) any = 1; if (c >= 4) exc.push_back((uint32_t)i); }
        int64_t i = lo;
#if defined(__x86_64__)
        if (simd == 2) i = dense_pack_avx2(codes + c0, c2, i, hi, xlo, xhi, exc, any);
        else if (simd == 1) i = dense_pack_sse41(codes + c0, c2, i, hi, xlo, xhi, exc, any);
#endif
        for (; i + 8 <= hi; i += 8) {              // 8 alleles -> 16 bits (code & 3: N, -, _ store code - 4)
            const uint64_t x = ld64(codes + c0 + i);
            any |= (x & 0xf8f8f8f8f8f8f8f8ull) | (x & (x >> 1) & (x >> 2) & 0x0101010101010101ull);   // a code > 6
            if (x & 0x0404040404040404ull) note(i, x);
            uint64_t y = x & 0x0303030303030303ull;
            y = (y | (y >> 6)) & 0x000f000f000f000full;
            y = (y | (y >> 12)) & 0x000000ff000000ffull;
            y = (y | (y >> 24)) & 0xffffull;
            const uint16_t v = (uint16_t)y;
            memcpy(c2 + (i >> 2), &v, 2);
        }
        for (; i < hi; i += 4) {
            uint8_t byte = 0;
            const int64_t m = std::min<int64_t>(4, n_codes - i);
            for (int64_t j = 0; j < m; ++j) {
                const uint8_t c = codes[c0 + i + j];
                if (c > 6) any = 1;
                if (c >= 4 && i + j >= xlo && i + j < xhi) exc.push_back((uint32_t)(i + j));
                byte |= (uint8_t)((c & 3) << (2 * j));
            }
            c2[i >> 2] = byte;
        }
        // (alleles in [hi, xhi) cannot exist: hi >= xhi by construction; alleles in [xhi, hi) are the next thread's
        // reads, whose exceptions it lists itself in its head loop)
        if (any) bd = 3;
        bad[(size_t)t] = bd;
    });
    int64_t tot_esc = 0, tot_exc = 0;
    for (int t = 0; t < nt; ++t) {
        if (bad[(size_t)t]) {
            hx_set_error("hx_dense_encode: %s", bad[(size_t)t] == 1 ? "reads are not sorted by rank"
                                               : bad[(size_t)t] == 2 ? "a read covers more SNPs than announced (or off[] decreases)"
                                                                     : "allele code > 6");
            return bad[(size_t)t] == 1 ? HX_E_STATE : HX_E_ARG;
        }
        tot_esc += (int64_t)P.esc_idx[t].size();
        tot_exc += (int64_t)P.exc[t].size();
    }
    P.n_esc = tot_esc; P.n_exc = tot_exc;
    P.o_esc_idx = P.o_exc + al16(tot_exc * 4);
    P.o_esc_delta = P.o_esc_idx + al16(tot_esc * 8);
    P.bytes = P.o_esc_delta + al16(tot_esc * 4) + 16;
    return HX_OK;
}

void hx_dense_finish(const HxDensePlan *pl, uint8_t *blob) {
    const HxDensePlan &P = *pl;
    uint32_t *exc = (uint32_t *)(blob + P.o_exc);
    int64_t *ei = (int64_t *)(blob + P.o_esc_idx);
    int32_t *ed = (int32_t *)(blob + P.o_esc_delta);
    for (int t = 0; t < P.nt; ++t) {
        if (!P.exc[t].empty()) memcpy(exc, P.exc[t].data(), 4 * P.exc[t].size());
        exc += P.exc[t].size();
        if (!P.esc_idx[t].empty()) {
            memcpy(ei, P.esc_idx[t].data(), 8 * P.esc_idx[t].size());
            memcpy(ed, P.esc_delta[t].data(), 4 * P.esc_delta[t].size());
        }
        ei += P.esc_idx[t].size();
        ed += P.esc_delta[t].size();
    }
}

extern "C" {

int hx_dense_encode(const int32_t *rank, const int64_t *off, const uint8_t *codes, int64_t n_reads, int n_threads,
                    hx_dense *out) {
    if (!out || n_reads < 0 || (n_reads > 0 && (!rank || !off || !codes))) {
        hx_set_error("hx_dense_encode: bad arguments");
        return HX_E_ARG;
    }
    memset(out, 0, sizeof(*out));
    HxDensePlan P;
    int rc = hx_dense_begin(off, n_reads, n_threads, 0, &P);
    if (rc) return rc;
    std::vector<uint8_t> head((size_t)P.head_bytes + 16, 0);
    rc = hx_dense_pack(rank, off, codes, &P, head.data());
    if (rc) return rc == HX_E_STATE ? HX_E_ARG : rc;
    uint8_t *blob = (uint8_t *)calloc(1, (size_t)P.bytes);
    if (!blob) { hx_set_error("hx_dense_encode: out of memory (%lld bytes)", (long long)P.bytes); return HX_E_NOMEM; }
    memcpy(blob, head.data(), (size_t)P.head_bytes);
    hx_dense_finish(&P, blob);
    out->blob = blob; out->blob_bytes = P.bytes; out->n_reads = n_reads; out->n_codes = P.n_codes;
    out->n_exc = P.n_exc; out->n_esc = P.n_esc; out->klen_bytes = P.klen_bytes;
    out->o_klen = P.o_klen; out->o_codes2 = P.o_codes2; out->o_exc = P.o_exc; out->o_esc_idx = P.o_esc_idx;
    out->o_esc_delta = P.o_esc_delta;
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
