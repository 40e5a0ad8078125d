// libtecbam: BAM (BGZF) -> structure-of-arrays batches on the host cores.  ABI: include/tecbam.h.
//
// Data flow per refill: the compressed file is mapped; a serial header walk finds the next run of
// BGZF blocks (block size from the BC extra field, uncompressed size from the trailer), so every
// block knows its place in one contiguous window before anything is inflated; the worker pool
// inflates the blocks in parallel (raw deflate + CRC32 check).  Record boundaries are a chain (each
// record starts where the previous one ends), so the task that has just inflated a run of blocks
// also hops over the block_size fields inside it while the bytes are still in its cache: it waits
// for the predecessor task to publish where its last record ended, lists the record starts of its
// own run and publishes the next start.  The chain costs a few ns per record instead of a cold
// miss per record in a serial pass.  The pool then parses the listed records in parallel straight
// into the caller's arrays.  A record cut by the end of the window is carried to the front of the
// next one.
//
// Field semantics follow te_counter_b200/reads.py + bam.py (the Python packing this replaces),
// which in turn cite the reference's read loops; see the header for file:line.
#include "../../include/tecbam.h"
#include "fast_inflate.h"

#include <zlib.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

constexpr size_t WINDOW_BYTES = size_t(48) << 20;      // uncompressed bytes inflated per refill (TEC_BAM_WINDOW overrides)
constexpr int PARSE_GRAIN = 2048;                       // records per parse task
constexpr int UMI_MAX_LEN = 21;
constexpr int F_UNMAPPED = 1, F_DUP = 2, F_QCFAIL = 4, F_REVERSE = 8, F_NAME_MISMATCH = 16;

inline uint16_t le16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }
inline uint32_t le32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline int32_t le32s(const uint8_t *p) { int32_t v; memcpy(&v, p, 4); return v; }

// ---------------------------------------------------------------------------------------------
// fork-join pool: run(n, f) calls f(i) for i in [0, n) on the workers and the calling thread
class Pool {
public:
    explicit Pool(int n_threads) {
        for (int i = 1; i < n_threads; i++) workers_.emplace_back([this] { loop(); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    int size() const { return int(workers_.size()) + 1; }

    void run(int n, const std::function<void(int)> &f) {
        if (n <= 0) return;
        if (workers_.empty() || n == 1) {
            for (int i = 0; i < n; i++) f(i);
            return;
        }
        {
            std::lock_guard<std::mutex> g(m_);
            fn_ = &f;
            n_ = n;
            next_.store(0);
            busy_ = int(workers_.size());
            gen_++;
        }
        cv_.notify_all();
        drain();
        std::unique_lock<std::mutex> g(m_);
        done_.wait(g, [this] { return busy_ == 0; });
        fn_ = nullptr;
    }

private:
    void drain() {
        for (;;) {
            int i = next_.fetch_add(1);
            if (i >= n_) break;
            (*fn_)(i);
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
            }
            drain();
            {
                std::lock_guard<std::mutex> g(m_);
                if (--busy_ == 0) done_.notify_one();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int)> *fn_ = nullptr;
    std::atomic<int> next_{0};
    int n_ = 0, busy_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

// ---------------------------------------------------------------------------------------------
struct Block {
    const uint8_t *cdata;
    uint32_t clen, isize, crc;
    size_t out;                 // offset in the window
};

struct Whitelist {
    std::string bytes;
    std::vector<int64_t> off;
    std::vector<uint32_t> slot;         // open addressing, value = id + 1
    uint64_t mask = 0;

    static uint64_t hash(const uint8_t *p, size_t n) {
        uint64_t h = 0xcbf29ce484222325ull;             // FNV-1a, then a finishing mix
        for (size_t i = 0; i < n; i++) h = (h ^ p[i]) * 0x100000001b3ull;
        h ^= h >> 29;
        h *= 0xbf58476d1ce4e5b9ull;
        return h ^ (h >> 32);
    }
    void build() {
        size_t n = off.size() - 1, cap = 16;
        while (cap < 2 * n + 2) cap <<= 1;
        slot.assign(cap, 0);
        mask = cap - 1;
        for (size_t i = 0; i < n; i++) {
            uint64_t h = hash((const uint8_t *)bytes.data() + off[i], size_t(off[i + 1] - off[i])) & mask;
            while (slot[h]) h = (h + 1) & mask;
            slot[h] = uint32_t(i + 1);
        }
    }
    uint32_t find(const uint8_t *p, size_t n) const {
        if (slot.empty()) return TBAM_CELL_INVALID;
        uint64_t h = hash(p, n) & mask;
        while (uint32_t v = slot[h]) {
            int64_t a = off[v - 1], b = off[v];
            if (size_t(b - a) == n && memcmp(bytes.data() + a, p, n) == 0) return v - 1;
            h = (h + 1) & mask;
        }
        return TBAM_CELL_INVALID;
    }
};

struct TaskError {
    int64_t rec = -1;
    int status = 0;
};

}  // namespace

struct tbam_reader {
    int fd = -1;
    const uint8_t *map = nullptr;
    size_t size = 0, pos = 0;                   // compressed file and cursor
    std::vector<uint8_t> win;                   // uncompressed window; [w_beg, w_end) is unread
    size_t w_beg = 0, w_end = 0;
    size_t window = WINDOW_BYTES;
    bool use_fast_inflate = true;               // TEC_BAM_INFLATE=zlib turns the table-driven decoder off
    std::vector<Block> blocks;
    std::vector<uint32_t> rec_off;              // starts of the complete records of the window
    size_t rec_cur = 0;                         // first one not handed out yet
    bool listed = false;                        // rec_off describes the current window
    std::vector<std::string> refs;
    std::vector<uint16_t> bulk_ids, sc_ids;
    int32_t n_index = 0;
    bool have_map = false;
    Whitelist wl;
    Pool *pool = nullptr;
    std::string err;
    int64_t n_records = 0, c_bytes = 0, u_bytes = 0, ns_next = 0;

    ~tbam_reader() {
        delete pool;
        if (map && size) munmap((void *)map, size);
        if (fd >= 0) close(fd);
    }
};

namespace {

int fail(tbam_reader *r, int status, const std::string &msg) {
    r->err = msg;
    return status;
}

// One BGZF block: the table-driven decoder of fast_inflate.h first; zlib decides whenever that one
// declines (it returns false on anything unusual), then the CRC32 of the trailer.
bool zlib_inflate(const uint8_t *in, uint32_t in_n, uint8_t *out, uint32_t out_n) {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) return false;
    zs.next_in = const_cast<Bytef *>(in);
    zs.avail_in = in_n;
    zs.next_out = out;
    zs.avail_out = out_n;
    int rc = inflate(&zs, Z_FINISH);
    bool ok = rc == Z_STREAM_END && zs.avail_out == 0;
    inflateEnd(&zs);
    return ok;
}

bool inflate_block(const Block &b, uint8_t *out, bool fast) {
    static thread_local fast_inflate::Inflater inf;
    if (!(fast && inf.run(b.cdata, b.clen, out, b.isize)) && !zlib_inflate(b.cdata, b.clen, out, b.isize)) return false;
    return uint32_t(crc32(crc32(0L, Z_NULL, 0), out, b.isize)) == b.crc;
}

// Lists the starts of the complete records in [from, w_end) by hopping over block_size fields.
// Returns the offset behind the last complete record, or SIZE_MAX for an impossible block_size.
size_t hop_records(const uint8_t *w, size_t from, size_t limit, size_t w_end, std::vector<uint32_t> &out) {
    size_t p = from;
    while (p + 4 <= limit) {
        uint32_t bs = le32(w + p);
        if (bs < 32 || bs > (1u << 30)) return SIZE_MAX;
        out.push_back(uint32_t(p));
        p += 4 + size_t(bs);
    }
    (void)w_end;
    return p;
}

// Finds the blocks of the next window and inflates them behind the carried-over bytes; with
// `list` also lists the record starts of the new window in rec_off (see the file comment).
// Leaves w_end unchanged when the compressed file is exhausted.
int refill(tbam_reader *r, bool list) {
    size_t keep = r->w_end - r->w_beg;
    if (r->w_beg && keep) memmove(r->win.data(), r->win.data() + r->w_beg, keep);
    r->w_beg = 0;
    r->w_end = keep;
    r->rec_off.clear();
    r->rec_cur = 0;
    r->listed = false;
    r->blocks.clear();
    size_t total = 0, p = r->pos;
    while (p < r->size && total < r->window) {
        if (r->size - p < 18) return fail(r, TBAM_E_FORMAT, "truncated BGZF block header");
        const uint8_t *h = r->map + p;
        if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8) return fail(r, TBAM_E_FORMAT, "bad gzip magic inside the file");
        if (!(h[3] & 4)) return fail(r, TBAM_E_NOT_BGZF, "gzip member without an extra field: not BGZF");
        uint32_t xlen = le16(h + 10);
        if (r->size - p < 12 + size_t(xlen)) return fail(r, TBAM_E_FORMAT, "truncated BGZF extra field");
        int64_t bsize = -1;
        for (uint32_t o = 0; o + 4 <= xlen;) {
            uint32_t slen = le16(h + 12 + o + 2);
            if (h[12 + o] == 'B' && h[12 + o + 1] == 'C' && slen == 2 && o + 6 <= xlen) bsize = int64_t(le16(h + 12 + o + 4)) + 1;
            o += 4 + slen;
        }
        if (bsize < 0) return fail(r, TBAM_E_NOT_BGZF, "gzip member without a BC field: not BGZF");
        if (bsize < int64_t(12 + xlen + 8) || size_t(bsize) > r->size - p)
            return fail(r, TBAM_E_FORMAT, "truncated BGZF block");
        Block b;
        b.cdata = h + 12 + xlen;
        b.clen = uint32_t(bsize - 12 - xlen - 8);
        b.crc = le32(h + bsize - 8);
        b.isize = le32(h + bsize - 4);
        if (b.isize > 65536) return fail(r, TBAM_E_FORMAT, "BGZF block larger than 64 KiB");
        b.out = keep + total;
        total += b.isize;
        p += size_t(bsize);
        if (b.isize) r->blocks.push_back(b);
    }
    r->c_bytes += int64_t(p - r->pos);
    r->pos = p;
    if (!total) return 0;
    if (r->win.size() < keep + total) r->win.resize(keep + total);
    uint8_t *base = r->win.data();
    const size_t w_end = keep + total;
    const std::vector<Block> &bl = r->blocks;
    const int per = 8;                                  // blocks per task
    const int n_tasks = int((bl.size() + per - 1) / per);
    std::atomic<int> bad{0};                            // 1 inflate / CRC, 2 record chain
    // chain[t] = start of the first record task t has to list (-1 until its predecessor knows);
    // task t lists the records that start before the end of its own bytes
    std::vector<std::atomic<int64_t>> chain(size_t(n_tasks) + 1);
    for (auto &c : chain) c.store(-1, std::memory_order_relaxed);
    chain[0].store(0, std::memory_order_release);
    std::vector<std::vector<uint32_t>> found(list ? size_t(n_tasks) : 0);
    r->pool->run(n_tasks, [&](int t) {
        size_t lo = size_t(t) * per, hi = std::min(bl.size(), lo + per);
        bool ok = true;
        for (size_t i = lo; ok && i < hi; i++) ok = inflate_block(bl[i], base + bl[i].out, r->use_fast_inflate);
        if (!ok) bad.store(1);
        if (!list) return;
        int64_t from;
        while ((from = chain[size_t(t)].load(std::memory_order_acquire)) < 0) std::this_thread::yield();
        size_t limit = bl[hi - 1].out + bl[hi - 1].isize, next = size_t(from);
        if (!ok || from == INT64_MAX) next = size_t(INT64_MAX);         // pass the failure on
        else if (size_t(from) + 4 <= limit) {
            found[size_t(t)].reserve((limit - size_t(from)) / 160 + 16);
            next = hop_records(base, size_t(from), limit, w_end, found[size_t(t)]);
            if (next == SIZE_MAX) {
                bad.store(2);
                next = size_t(INT64_MAX);
            }
        }
        chain[size_t(t) + 1].store(int64_t(next), std::memory_order_release);
    });
    if (bad.load() == 1) return fail(r, TBAM_E_FORMAT, "corrupt BGZF block (inflate or CRC32 failed)");
    if (bad.load() == 2) return fail(r, TBAM_E_FORMAT, "alignment record with impossible block_size");
    r->w_end = w_end;
    r->u_bytes += int64_t(total);
    if (list) {
        size_t n = 0;
        for (auto &v : found) n += v.size();
        r->rec_off.reserve(n);
        for (auto &v : found) r->rec_off.insert(r->rec_off.end(), v.begin(), v.end());
        // only the last listed record can reach beyond the window: it is the carry
        if (!r->rec_off.empty() && size_t(chain[size_t(n_tasks)].load()) > w_end) r->rec_off.pop_back();
        r->listed = true;
    }
    return 0;
}

// Lists the records of a window that was inflated without listing (the one holding the header).
int list_window(tbam_reader *r) {
    r->rec_off.clear();
    r->rec_cur = 0;
    size_t next = hop_records(r->win.data(), r->w_beg, r->w_end, r->w_end, r->rec_off);
    if (next == SIZE_MAX) return fail(r, TBAM_E_FORMAT, "alignment record with impossible block_size");
    if (!r->rec_off.empty() && next > r->w_end) r->rec_off.pop_back();
    r->listed = true;
    return 0;
}

// Makes at least n unread bytes available; returns 1 when the stream ends first.
int need(tbam_reader *r, size_t n) {
    while (r->w_end - r->w_beg < n) {
        size_t before = r->w_end - r->w_beg;
        int rc = refill(r, false);
        if (rc) return rc;
        if (r->w_end - r->w_beg == before && r->pos >= r->size) return 1;
    }
    return 0;
}

int read_header(tbam_reader *r) {
    int rc = need(r, 12);
    if (rc < 0) return rc;
    if (rc || memcmp(r->win.data() + r->w_beg, "BAM\1", 4) != 0) return fail(r, TBAM_E_FORMAT, "not a BAM file (magic)");
    int32_t l_text = le32s(r->win.data() + r->w_beg + 4);
    if (l_text < 0) return fail(r, TBAM_E_FORMAT, "negative header text length");
    rc = need(r, 12 + size_t(l_text));
    if (rc) return rc < 0 ? rc : fail(r, TBAM_E_FORMAT, "truncated BAM header");
    r->w_beg += 8 + size_t(l_text);
    int32_t n_ref = le32s(r->win.data() + r->w_beg);
    r->w_beg += 4;
    if (n_ref < 0) return fail(r, TBAM_E_FORMAT, "negative reference count");
    for (int32_t i = 0; i < n_ref; i++) {
        rc = need(r, 4);
        if (rc) return rc < 0 ? rc : fail(r, TBAM_E_FORMAT, "truncated BAM reference list");
        int32_t l_name = le32s(r->win.data() + r->w_beg);
        if (l_name < 0) return fail(r, TBAM_E_FORMAT, "negative reference name length");
        rc = need(r, 8 + size_t(l_name));
        if (rc) return rc < 0 ? rc : fail(r, TBAM_E_FORMAT, "truncated BAM reference list");
        const char *nm = (const char *)r->win.data() + r->w_beg + 4;
        r->refs.emplace_back(nm, strnlen(nm, size_t(l_name)));
        r->w_beg += 8 + size_t(l_name);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// one alignment record; p points behind block_size, n = block_size
struct Rec {
    int32_t ref_id, pos, end;       // end = -1 when has_end is false (pysam's None)
    bool has_end;
    uint8_t mapq, fbits;
    const uint8_t *name;
    uint32_t l_name;                // without the NUL
    const uint8_t *aux, *aux_end;
    bool ok;
};

inline Rec parse_core(const uint8_t *p, uint32_t n) {
    Rec c;
    c.ref_id = le32s(p);
    c.pos = le32s(p + 4);
    uint32_t l_name = p[8];
    c.mapq = p[9];
    uint32_t n_cig = le16(p + 12), flag = le16(p + 14);
    int64_t l_seq = le32s(p + 16);
    c.fbits = uint8_t(((flag >> 2) & 1) * F_UNMAPPED | ((flag >> 10) & 1) * F_DUP | ((flag >> 9) & 1) * F_QCFAIL |
                      ((flag >> 4) & 1) * F_REVERSE);
    size_t o = 32 + size_t(l_name) + 4 * size_t(n_cig);
    c.ok = l_seq >= 0 && l_name >= 1 && o + size_t((l_seq + 1) / 2) + size_t(l_seq) <= n;
    if (!c.ok) return c;
    c.name = p + 32;
    c.l_name = uint32_t(strnlen((const char *)c.name, l_name - 1));
    c.end = -1;
    c.has_end = n_cig && !(flag & 4);
    if (c.has_end) {                 // pysam: reference_end is None without a CIGAR / when unmapped
        const uint8_t *cg = p + 32 + l_name;
        uint32_t span = 0;
        for (uint32_t i = 0; i < n_cig; i++) {
            uint32_t v = le32(cg + 4 * i), op = v & 15;
            if ((0x18Du >> op) & 1) span += v >> 4;     // M D N = X consume the reference
        }
        c.end = int32_t(uint32_t(c.pos) + (span ? span : 1u));      // htslib bam_endpos: pos + 1 when the CIGAR consumes no reference base
    }
    c.aux = p + o + size_t((l_seq + 1) / 2) + size_t(l_seq);
    c.aux_end = p + n;
    return c;
}

// '_'.join(name.split('/')[0:-1]): the bytes before the last '/', with '/' read as '_'
inline bool mate_names_match(const Rec &a, const Rec &b) {
    auto prefix = [](const Rec &r) {
        uint32_t k = r.l_name;
        while (k && r.name[k - 1] != '/') k--;
        return k ? k - 1 : 0u;
    };
    uint32_t la = prefix(a), lb = prefix(b);
    if (la != lb) return false;
    for (uint32_t i = 0; i < la; i++) {
        uint8_t x = a.name[i] == '/' ? '_' : a.name[i], y = b.name[i] == '/' ? '_' : b.name[i];
        if (x != y) return false;
    }
    return true;
}

struct TagVal {
    const uint8_t *p = nullptr;
    uint32_t n = 0;
    int kind = 0;               // 0 absent, 1 string (Z, H, A), 2 other type
};

// Walks the aux fields once; the last occurrence of a tag wins, as in dict(read.get_tags()).
// Returns false on a malformed aux block.
bool find_tags(const Rec &c, TagVal &cb, TagVal &cr, TagVal &ub, TagVal &ur) {
    const uint8_t *p = c.aux, *e = c.aux_end;
    while (p + 3 <= e) {
        uint8_t t0 = p[0], t1 = p[1], ty = p[2];
        p += 3;
        TagVal v;
        size_t adv;
        switch (ty) {
        case 'A': case 'c': case 'C': adv = 1; break;
        case 's': case 'S': adv = 2; break;
        case 'i': case 'I': case 'f': adv = 4; break;
        case 'Z': case 'H': {
            const uint8_t *z = (const uint8_t *)memchr(p, 0, size_t(e - p));
            if (!z) return false;
            adv = size_t(z - p) + 1;
            break;
        }
        case 'B': {
            if (p + 5 > e) return false;
            size_t sz;
            switch (p[0]) {
            case 'c': case 'C': sz = 1; break;
            case 's': case 'S': sz = 2; break;
            case 'i': case 'I': case 'f': sz = 4; break;
            default: return false;
            }
            adv = 5 + sz * size_t(le32(p + 1));
            break;
        }
        default: return false;
        }
        if (adv > size_t(e - p)) return false;
        if ((t0 == 'C' || t0 == 'U') && (t1 == 'B' || t1 == 'R')) {
            v.p = p;
            if (ty == 'Z' || ty == 'H') { v.n = uint32_t(adv - 1); v.kind = 1; }
            else if (ty == 'A') { v.n = 1; v.kind = 1; }
            else v.kind = 2;
            (t0 == 'C' ? (t1 == 'B' ? cb : cr) : (t1 == 'B' ? ub : ur)) = v;
        }
        p += adv;
    }
    return true;                // a tail shorter than one field is ignored, as in bam.py
}

// reads.encode_umi: A1 C2 G3 N4 T5, 3 bits per character, left aligned in 63 bits
inline bool encode_umi(const uint8_t *s, uint32_t n, uint64_t &code) {
    if (n > UMI_MAX_LEN) return false;
    uint64_t c = 0;
    for (uint32_t i = 0; i < n; i++) {
        uint64_t v;
        switch (s[i]) {
        case 'A': v = 1; break;
        case 'C': v = 2; break;
        case 'G': v = 3; break;
        case 'N': v = 4; break;
        case 'T': v = 5; break;
        default: return false;
        }
        c = (c << 3) | v;
    }
    code = c << (3 * (UMI_MAX_LEN - n));
    return true;
}

struct Out {
    int32_t *start, *end;
    uint16_t *chrom;
    uint8_t *mapq, *flag;
    uint32_t *cell;
    uint64_t *umi;
};

enum Mode { MODE_SE, MODE_PE, MODE_SC };

// Parses records [a, b) of rec_off into out[base + a ...]; returns the first failing record.
TaskError parse_range(const tbam_reader *r, Mode mode, int qual, const Out &o, int64_t base, size_t a, size_t b) {
    TaskError te;
    const uint8_t *w = r->win.data();
    const int32_t n_ref = int32_t(r->refs.size());
    auto core = [&](size_t i, Rec &c) {
        const uint8_t *p = w + r->rec_off[i];
        c = parse_core(p + 4, le32(p));
        return c.ok;
    };
    auto bad = [&](size_t i, int status) {
        te.rec = int64_t(i);
        te.status = status;
        return te;
    };
    auto bulk_chrom = [&](const Rec &c) { return c.ref_id >= 0 && c.ref_id < n_ref ? r->bulk_ids[c.ref_id] : uint16_t(TBAM_CHROM_INVALID); };
    if (mode == MODE_PE) {
        for (size_t i = a; i + 1 < b; i += 2) {
            Rec c1, c2;
            if (!core(i, c1)) return bad(i, TBAM_E_FORMAT);
            if (!core(i + 1, c2)) return bad(i + 1, TBAM_E_FORMAT);
            uint8_t f1 = c1.fbits;
            bool rejected = ((c1.fbits | c2.fbits) & (F_UNMAPPED | F_DUP | F_QCFAIL)) || int(c1.mapq) < qual;
            if (!rejected && !mate_names_match(c1, c2)) f1 |= F_NAME_MISMATCH;
            int64_t k = base + int64_t(i);
            o.start[k] = c1.pos; o.end[k] = c1.end; o.chrom[k] = bulk_chrom(c1); o.mapq[k] = c1.mapq; o.flag[k] = f1;
            k++;
            o.start[k] = c2.pos; o.end[k] = c2.end; o.chrom[k] = bulk_chrom(c2); o.mapq[k] = c2.mapq; o.flag[k] = c2.fbits;
        }
        return te;
    }
    if (mode == MODE_SE) {
        for (size_t i = a; i < b; i++) {
            Rec c;
            if (!core(i, c)) return bad(i, TBAM_E_FORMAT);
            uint16_t ch = bulk_chrom(c);
            if (!c.has_end && !(c.fbits & (F_UNMAPPED | F_DUP | F_QCFAIL)) && int(c.mapq) >= qual && int32_t(ch) < r->n_index)
                return bad(i, TBAM_E_END_NONE);                 // te_count.py:223, (loc2+1) with loc2 None
            int64_t k = base + int64_t(i);
            o.start[k] = c.pos; o.end[k] = c.end; o.chrom[k] = ch; o.mapq[k] = c.mapq; o.flag[k] = c.fbits;
        }
        return te;
    }
    for (size_t i = a; i < b; i++) {
        Rec c;
        if (!core(i, c)) return bad(i, TBAM_E_FORMAT);
        int64_t k = base + int64_t(i);
        o.mapq[k] = c.mapq;
        o.flag[k] = c.fbits;
        o.start[k] = -1; o.end[k] = -1; o.chrom[k] = uint16_t(TBAM_CHROM_INVALID); o.cell[k] = TBAM_CELL_INVALID; o.umi[k] = 0;
        if ((c.fbits & (F_UNMAPPED | F_DUP | F_QCFAIL)) || int(c.mapq) < qual) continue;
        TagVal cb, cr, ub, ur;
        if (!find_tags(c, cb, cr, ub, ur)) return bad(i, TBAM_E_FORMAT);
        const TagVal &bc = cb.kind ? cb : cr;
        if (!bc.kind) return bad(i, TBAM_E_NO_BARCODE_TAG);
        uint32_t cid = bc.kind == 1 ? r->wl.find(bc.p, bc.n) : TBAM_CELL_INVALID;
        if (cid == TBAM_CELL_INVALID) continue;                 // te_count.py:412 invalid barcode
        const TagVal &um = ub.kind ? ub : ur;
        if (!um.kind) return bad(i, TBAM_E_NO_UMI_TAG);
        uint64_t code;
        if (um.kind != 1 || !encode_umi(um.p, um.n, code)) return bad(i, TBAM_E_UMI);
        if (c.ref_id < 0 || c.ref_id >= n_ref) return bad(i, TBAM_E_REF_NONE);
        uint16_t ch = r->sc_ids[c.ref_id];
        if (ch == TBAM_CHROM_SC_BAD) return bad(i, TBAM_E_CHROM_NAME);
        if (!c.has_end && ch != TBAM_CHROM_SC_SKIP) return bad(i, TBAM_E_END_NONE);
        o.start[k] = c.pos; o.end[k] = c.end; o.chrom[k] = ch; o.cell[k] = cid; o.umi[k] = code;
    }
    return te;
}

const char *status_text(int s) {
    switch (s) {
    case TBAM_OK: return "ok";
    case TBAM_E_IO: return "cannot open or map the file";
    case TBAM_E_FORMAT: return "malformed BAM data";
    case TBAM_E_NOT_BGZF: return "not a BGZF file";
    case TBAM_E_ARG: return "bad argument";
    case TBAM_E_NO_BARCODE_TAG: return "CB or CR tag not found!";
    case TBAM_E_NO_UMI_TAG: return "UB or UR tag not found!";
    case TBAM_E_UMI: return "UMI longer than 21 characters or with a character outside A,C,G,N,T";
    case TBAM_E_END_NONE: return "reference_end is None for a counted read";
    case TBAM_E_CHROM_NAME: return "chromosome name contains ':' (unsupported in --sc)";
    case TBAM_E_REF_NONE: return "record without a reference sequence passed the filters";
    }
    return "unknown status";
}

// End of the record that starts at window offset p.
inline size_t rec_end(const tbam_reader *r, size_t p) { return p + 4 + size_t(le32(r->win.data() + p)); }

int next_batch(tbam_reader *r, Mode mode, int qual, int64_t capacity, const Out &o, int64_t *n_out, int *more) {
    if (!r || !n_out || !more || capacity < 0 || !o.start || !o.end || !o.chrom || !o.mapq || !o.flag ||
        (mode == MODE_SC && (!o.cell || !o.umi)))
        return r ? fail(r, TBAM_E_ARG, "null buffer or negative capacity") : TBAM_E_ARG;
    if (!r->have_map) return fail(r, TBAM_E_ARG, "tbam_set_chrom_map has not been called");
    auto t0 = std::chrono::steady_clock::now();
    const bool paired = mode == MODE_PE;
    int64_t cap = paired ? capacity & ~int64_t(1) : capacity, n = 0;
    *more = 1;
    *n_out = 0;
    while (n < cap) {
        if (!r->listed) {
            int rc = list_window(r);
            if (rc) return rc;
        }
        size_t k = std::min(size_t(cap - n), r->rec_off.size() - r->rec_cur);
        if (paired) k &= ~size_t(1);                    // pairs never straddle two parse rounds
        if (k == 0) {
            size_t before = r->w_end - r->w_beg;
            int rc = refill(r, true);
            if (rc) return rc;
            if (r->w_end - r->w_beg == before) {        // no new bytes
                if (r->pos < r->size) continue;         // (only empty blocks so far)
                bool lone_mate = paired && before >= 36 && rec_end(r, r->w_beg) == r->w_end;
                if (before && !lone_mate) return fail(r, TBAM_E_FORMAT, "truncated BAM file (partial record at the end)");
                r->w_beg = r->w_end;                    // a trailing unpaired record is dropped (te_count.py:79)
                r->rec_off.clear();
                r->rec_cur = 0;
                *more = 0;
                break;
            }
            continue;
        }
        const size_t first = r->rec_cur;
        int n_tasks = int((k + PARSE_GRAIN - 1) / PARSE_GRAIN);
        std::vector<TaskError> errs(static_cast<size_t>(n_tasks), TaskError{});
        r->pool->run(n_tasks, [&](int t) {
            size_t a = size_t(t) * PARSE_GRAIN, b = std::min(k, a + PARSE_GRAIN);
            errs[size_t(t)] = parse_range(r, mode, qual, o, n - int64_t(first), first + a, first + b);
        });
        for (const TaskError &e : errs)
            if (e.status) {
                char buf[160];
                snprintf(buf, sizeof buf, "%s (record %lld)", status_text(e.status),
                         (long long)(r->n_records + n + e.rec - int64_t(first)));
                return fail(r, e.status, buf);
            }
        r->rec_cur += k;
        r->w_beg = rec_end(r, r->rec_off[r->rec_cur - 1]);
        n += int64_t(k);
    }
    r->n_records += n;
    *n_out = n;
    r->ns_next += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    return TBAM_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
extern "C" {

int tbam_abi_version(void) { return TBAM_ABI_VERSION; }

const char *tbam_strerror(int status) { return status_text(status); }

int tbam_open(const char *path, int n_threads, tbam_reader **out) {
    if (!path || !out) return TBAM_E_ARG;
    *out = nullptr;
    tbam_reader *r = new tbam_reader();
    r->fd = open(path, O_RDONLY);
    struct stat st;
    if (r->fd < 0 || fstat(r->fd, &st) != 0) {
        delete r;
        return TBAM_E_IO;
    }
    r->size = size_t(st.st_size);
    if (r->size < 28) {
        delete r;
        return TBAM_E_NOT_BGZF;
    }
    void *m = mmap(nullptr, r->size, PROT_READ, MAP_PRIVATE, r->fd, 0);
    if (m == MAP_FAILED) {
        r->size = 0;
        delete r;
        return TBAM_E_IO;
    }
    r->map = (const uint8_t *)m;
    madvise(m, r->size, MADV_SEQUENTIAL);
    if (r->map[0] != 0x1f || r->map[1] != 0x8b) {
        delete r;
        return TBAM_E_NOT_BGZF;
    }
    if (n_threads <= 0) {
        n_threads = int(std::thread::hardware_concurrency());
        if (n_threads <= 0) n_threads = 1;
        if (n_threads > 64) n_threads = 64;
    }
    r->pool = new Pool(n_threads);
    if (const char *z = getenv("TEC_BAM_INFLATE")) r->use_fast_inflate = strcmp(z, "zlib") != 0;
    if (const char *w = getenv("TEC_BAM_WINDOW")) {
        long long v = atoll(w);
        if (v > 0) r->window = size_t(v);
    }
    int rc = read_header(r);
    if (rc) {
        delete r;
        return rc;
    }
    *out = r;
    return TBAM_OK;
}

void tbam_close(tbam_reader *r) { delete r; }

const char *tbam_last_error(const tbam_reader *r) { return r ? r->err.c_str() : "null reader"; }

int tbam_n_references(const tbam_reader *r) { return r ? int(r->refs.size()) : 0; }

const char *tbam_reference_name(const tbam_reader *r, int i) {
    return r && i >= 0 && size_t(i) < r->refs.size() ? r->refs[size_t(i)].c_str() : nullptr;
}

int tbam_set_chrom_map(tbam_reader *r, const uint16_t *bulk_ids, const uint16_t *sc_ids, int32_t n, int32_t n_index) {
    if (!r) return TBAM_E_ARG;
    if (n != int32_t(r->refs.size()) || n_index < 0 || (n && (!bulk_ids || !sc_ids)))
        return fail(r, TBAM_E_ARG, "chrom map must have one entry per reference sequence");
    r->bulk_ids.assign(bulk_ids, bulk_ids + n);
    r->sc_ids.assign(sc_ids, sc_ids + n);
    r->n_index = n_index;
    r->have_map = true;
    return TBAM_OK;
}

int tbam_set_whitelist(tbam_reader *r, const char *barcodes, const int64_t *offsets, int32_t n) {
    if (!r) return TBAM_E_ARG;
    if (n < 0 || !offsets || (!barcodes && n && offsets[n] > 0)) return fail(r, TBAM_E_ARG, "bad whitelist arguments");
    for (int32_t i = 0; i < n; i++)
        if (offsets[i + 1] < offsets[i] || offsets[0] != 0) return fail(r, TBAM_E_ARG, "whitelist offsets must ascend from 0");
    r->wl.off.assign(offsets, offsets + n + 1);
    r->wl.bytes.assign(barcodes ? barcodes : "", size_t(offsets[n]));
    r->wl.build();
    return TBAM_OK;
}

int tbam_next_bulk(tbam_reader *r, int paired, int qual, int64_t capacity, int32_t *start, int32_t *end,
                   uint16_t *chrom, uint8_t *mapq, uint8_t *flag, int64_t *n_out, int *more) {
    Out o{start, end, chrom, mapq, flag, nullptr, nullptr};
    return next_batch(r, paired ? MODE_PE : MODE_SE, qual, capacity, o, n_out, more);
}

int tbam_next_sc(tbam_reader *r, int qual, int64_t capacity, int32_t *start, int32_t *end, uint16_t *chrom,
                 uint8_t *mapq, uint8_t *flag, uint32_t *cell, uint64_t *umi, int64_t *n_out, int *more) {
    Out o{start, end, chrom, mapq, flag, cell, umi};
    return next_batch(r, MODE_SC, qual, capacity, o, n_out, more);
}

int tbam_inflate_raw(const void *in, int64_t n_in, void *out, int64_t n_out, int engine) {
    if (n_in < 0 || n_out < 0 || n_in > (int64_t(1) << 31) || n_out > (int64_t(1) << 31) || (n_in && !in) || (n_out && !out)) return TBAM_E_ARG;
    uint8_t dummy = 0;
    const uint8_t *i = in ? (const uint8_t *)in : &dummy;
    uint8_t *o = out ? (uint8_t *)out : &dummy;
    bool ok;
    if (engine == 1) {
        static thread_local fast_inflate::Inflater inf;
        ok = inf.run(i, size_t(n_in), o, size_t(n_out));
    } else {
        ok = zlib_inflate(i, uint32_t(n_in), o, uint32_t(n_out));
    }
    return ok ? TBAM_OK : TBAM_E_FORMAT;
}

int64_t tbam_counter(const tbam_reader *r, int what) {
    if (!r) return 0;
    switch (what) {
    case 0: return r->n_records;
    case 1: return r->c_bytes;
    case 2: return r->u_bytes;
    case 3: return r->pool ? r->pool->size() : 0;
    case 4: return r->ns_next;
    }
    return 0;
}

}  // extern "C"
