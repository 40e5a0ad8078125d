// TEST INFRASTRUCTURE: the block-parallel BAM decoder (te_counter_b200/csrc/bgzf_dev.h per-block
// routines + bam_orch.h window loop) with a backend of plain host loops in place of the CUDA
// kernels of bamgpu.cuh.  Same routines, same orchestration, one "thread" after the other, so the
// CPU-only build box can hold the whole logic against libtecbam (tests/test_bgzf_dev_cpu.py).
// Build: g++ -O2 -std=c++17 -shared -fPIC tools/bgzf_dev_host.cpp -lz -o tools/libbgzfdevhost.so
#include "../te_counter_b200/csrc/bam_orch.h"

#include <algorithm>

namespace {

struct HostBackend {
    std::vector<uint8_t> comp, ubuf, scratch;
    std::vector<int32_t> c_start, c_end;
    std::vector<uint16_t> c_chrom;
    std::vector<uint8_t> c_mapq, c_flag;
    std::vector<uint32_t> c_cell;
    std::vector<uint64_t> c_umi;
    // everything delivered so far
    std::vector<int32_t> o_start, o_end;
    std::vector<uint16_t> o_chrom;
    std::vector<uint8_t> o_mapq, o_flag;
    std::vector<uint32_t> o_cell;
    std::vector<uint64_t> o_umi;
    uint32_t crc_table[256], crc_mats[bgzfdev::CRC_SHIFT_MATS * 32];
    int n_inflate_declined = 0;
    int force_decline_every = 0;                // test hook: pretend the block-parallel inflate declined every k-th block

    HostBackend() {
        bamorch::crc32_tables(crc_table, crc_mats);
        scratch.resize(bgzfdev::SCRATCH_STRIDE);
    }
    int reserve(size_t comp_bytes, size_t ubuf_bytes, int) {
        if (comp.size() < comp_bytes) comp.resize(comp_bytes);
        if (ubuf.size() < ubuf_bytes) ubuf.resize(ubuf_bytes);
        return 0;
    }
    int load(const bamorch::MappedFile& f, size_t lo, size_t n) {
        memcpy(comp.data(), f.map + lo, n);
        return 0;
    }
    int put(int64_t at, const uint8_t* data, size_t n) {
        memcpy(ubuf.data() + at, data, n);
        return 0;
    }
    int carry(int64_t from, int64_t n) {
        memmove(ubuf.data(), ubuf.data() + from, (size_t)n);
        return 0;
    }
    int inflate(const bamorch::BlockDesc* bl, int nb, int32_t* status) {
        for (int b = 0; b < nb; b++) {          // one CUDA thread per b
            const bamorch::BlockDesc& d = bl[b];
            int st = bgzfdev::inflate_block(comp.data() + d.in_off, d.in_len, ubuf.data() + d.out_off, d.out_len, scratch.data());
            if (st == bgzfdev::ST_OK) {         // the device computes it with the 32 lanes of the warp: same shares here, one after the other
                uint32_t c = 0;
                for (int lane = 0; lane < 32; lane++) c ^= bgzfdev::crc32_lane_share(ubuf.data() + d.out_off, d.out_len, lane, crc_table, crc_mats);
                if ((c ^ 0xFFFFFFFFu) != d.crc || bgzfdev::crc32_block(ubuf.data() + d.out_off, d.out_len, crc_table) != d.crc) st = bgzfdev::ST_CRC;
            }
            if (force_decline_every && b % force_decline_every == 0) {
                memset(ubuf.data() + d.out_off, 0xAB, d.out_len);
                st = bgzfdev::ST_DECLINED;
            }
            n_inflate_declined += st != bgzfdev::ST_OK;
            status[b] = st;
        }
        return 0;
    }
    int chain(const bamorch::BlockDesc* bl, int nb, int64_t w_end, int32_t n_ref, bamorch::BlockChain* out) {
        for (int b = 0; b < nb; b++) {
            const int64_t lo = (int64_t)bl[b].out_off, hi = lo + bl[b].out_len;
            const int64_t s = b == 0 ? 0 : bgzfdev::find_start(ubuf.data(), lo, hi, w_end, n_ref);
            bamorch::BlockChain c;
            c.start = s; c.exit = s; c.last = bgzfdev::NO_START; c.count = 0; c.bad = 0;
            if (s != bgzfdev::NO_START) {
                const bgzfdev::Hop h = bgzfdev::hop(ubuf.data(), s, hi, w_end);
                c.exit = h.exit; c.last = h.last; c.count = h.count; c.bad = h.bad;
            }
            out[b] = c;
        }
        return 0;
    }
    int parse(const bamorch::BlockDesc* bl, int nb, const bamorch::BlockChain* ch, const int64_t* base, int64_t n, int64_t, int mode, int qual,
              const bamorch::Reader& r, int* err, int64_t* err_rec) {
        c_start.resize((size_t)n); c_end.resize((size_t)n); c_chrom.resize((size_t)n); c_mapq.resize((size_t)n); c_flag.resize((size_t)n);
        c_cell.resize((size_t)n); c_umi.resize((size_t)n);
        bgzfdev::Columns o{c_start.data(), c_end.data(), c_chrom.data(), c_mapq.data(), c_flag.data(), c_cell.data(), c_umi.data()};
        bgzfdev::ParseCtx pc;
        pc.bulk_ids = r.bulk_ids.data(); pc.sc_ids = r.sc_ids.data(); pc.n_ref = (int32_t)r.refs.size(); pc.n_index = r.n_index; pc.qual = qual;
        pc.wl.slot = r.wl_slot.data(); pc.wl.off = r.wl_off.data(); pc.wl.bytes = (const uint8_t*)r.wl_bytes.data();
        pc.wl.mask = r.wl_slot.empty() ? 0 : r.wl_slot.size() - 1;
        uint64_t first_err = ~uint64_t(0);
        (void)bl;
        for (int b = 0; b < nb; b++) {          // one CUDA thread per b
            int64_t p = ch[b].start;
            for (uint32_t i = 0; i < ch[b].count; i++) {
                const int64_t k = base[b] + i, nxt = p + 4 + (int64_t)bgzfdev::ld32(ubuf.data() + p);
                if (mode != bgzfdev::MODE_PE || !(k & 1)) {
                    const int e = bgzfdev::parse_record(ubuf.data(), p, nxt, mode, pc, o, k);
                    if (e) first_err = std::min(first_err, ((uint64_t)k << 8) | (uint64_t)e);
                }
                p = nxt;
            }
        }
        *err = 0;
        if (first_err != ~uint64_t(0)) {
            *err = (int)(first_err & 0xFF);
            *err_rec = (int64_t)(first_err >> 8);
        }
        return 0;
    }
    int deliver(int64_t n, int mode) {
        o_start.insert(o_start.end(), c_start.begin(), c_start.begin() + n);
        o_end.insert(o_end.end(), c_end.begin(), c_end.begin() + n);
        o_chrom.insert(o_chrom.end(), c_chrom.begin(), c_chrom.begin() + n);
        o_mapq.insert(o_mapq.end(), c_mapq.begin(), c_mapq.begin() + n);
        o_flag.insert(o_flag.end(), c_flag.begin(), c_flag.begin() + n);
        if (mode == bgzfdev::MODE_SC) {
            o_cell.insert(o_cell.end(), c_cell.begin(), c_cell.begin() + n);
            o_umi.insert(o_umi.end(), c_umi.begin(), c_umi.begin() + n);
        }
        return 0;
    }
};

struct Session {
    bamorch::Reader reader;
    HostBackend be;
};

}  // namespace

extern "C" {

int bgzfdev_open(const char* path, void** out) {
    Session* s = new Session();
    int rc = s->reader.open_path(path);
    if (rc) {
        delete s;
        return rc;
    }
    *out = s;
    return 0;
}
void bgzfdev_close(void* h) { delete (Session*)h; }
const char* bgzfdev_error(void* h) { return ((Session*)h)->reader.err.c_str(); }
int bgzfdev_n_references(void* h) { return (int)((Session*)h)->reader.refs.size(); }
const char* bgzfdev_reference_name(void* h, int i) { return ((Session*)h)->reader.refs[(size_t)i].c_str(); }
int bgzfdev_set_chrom_map(void* h, const uint16_t* b, const uint16_t* s, int32_t n, int32_t n_index) {
    return ((Session*)h)->reader.set_chrom_map(b, s, n, n_index);
}
int bgzfdev_set_whitelist(void* h, const char* bytes, const int64_t* off, int32_t n) { return ((Session*)h)->reader.set_whitelist(bytes, off, n); }
void bgzfdev_force_decline(void* h, int every) { ((Session*)h)->be.force_decline_every = every; }
int bgzfdev_declined(void* h) { return ((Session*)h)->be.n_inflate_declined; }

int bgzfdev_decode(void* h, int mode, int qual, int window_blocks, int64_t* n_records) {
    Session* s = (Session*)h;
    return bamorch::decode_all(s->reader, s->be, mode, qual, window_blocks, n_records);
}

// one byte range of the file (bam_orch.h decode_range); out[6] = {n_records, start_block, start_off, exit_block, exit_off, file size}
int bgzfdev_decode_range(void* h, int mode, int qual, int window_blocks, int64_t lo, int64_t hi, int64_t* out) {
    Session* s = (Session*)h;
    bamorch::RangeResult r;
    const int rc = bamorch::decode_range(s->reader, s->be, mode, qual, window_blocks, (uint64_t)lo, (uint64_t)hi, &r);
    out[0] = r.n_records; out[1] = r.start_block; out[2] = r.start_off; out[3] = r.exit_block; out[4] = r.exit_off;
    out[5] = (int64_t)s->reader.file.size;
    return rc;
}

// copies the delivered columns out (cell / umi only for single cell)
void bgzfdev_fetch(void* h, int32_t* start, int32_t* end, uint16_t* chrom, uint8_t* mapq, uint8_t* flag, uint32_t* cell, uint64_t* umi) {
    HostBackend& b = ((Session*)h)->be;
    const size_t n = b.o_start.size();
    if (!n) return;
    memcpy(start, b.o_start.data(), n * 4);
    memcpy(end, b.o_end.data(), n * 4);
    memcpy(chrom, b.o_chrom.data(), n * 2);
    memcpy(mapq, b.o_mapq.data(), n);
    memcpy(flag, b.o_flag.data(), n);
    if (cell && !b.o_cell.empty()) memcpy(cell, b.o_cell.data(), n * 4);
    if (umi && !b.o_umi.empty()) memcpy(umi, b.o_umi.data(), n * 8);
}

// CRC32 the way the warp computes it (32 lane shares), for holding it against zlib.crc32
uint32_t bgzfdev_crc32_lanes(const void* p, int64_t n) {
    static uint32_t table[256], mats[bgzfdev::CRC_SHIFT_MATS * 32];
    static bool init = false;
    if (!init) {
        bamorch::crc32_tables(table, mats);
        init = true;
    }
    uint8_t dummy = 0;
    uint32_t c = 0;
    for (int lane = 0; lane < 32; lane++) c ^= bgzfdev::crc32_lane_share(p ? (const uint8_t*)p : &dummy, (uint32_t)n, lane, table, mats);
    return c ^ 0xFFFFFFFFu;
}

// the block inflate alone, for holding it against zlib
int bgzfdev_inflate_raw(const void* in, int64_t n_in, void* out, int64_t n_out) {
    std::vector<uint8_t> scratch(bgzfdev::SCRATCH_STRIDE);
    uint8_t dummy = 0;
    return bgzfdev::inflate_block(in ? (const uint8_t*)in : &dummy, (uint32_t)n_in, out ? (uint8_t*)out : &dummy, (uint32_t)n_out, scratch.data());
}

}  // extern "C"
