// Single cell over several GPUs with the collectives issued by the library (collective.cuh): the exchange of the
// survivors by owner rank ("partition by cell id after filtering, all-to-all of the records keyed by
// cell, original record index as tiebreak") and the final all-gather + merge of the (ensg, cell, count) triples.
// te_counter_b200/dist.py keeps a torch.distributed / callback version of the same steps for the CPU-side tests (gloo).
#pragma once
#include "sc.cuh"
#include "collective.cuh"

// After the pushes of every rank: survivors packed by owner rank (cell % world), exchanged with grouped NCCL
// send / receive, installed as this rank's survivors (ascending in job-wide position).  Every rank must call it.
extern "C" int tec_sc_exchange(tec_ctx* ctx, int64_t* n_owned) {
    if (!ctx) return TEC_ERR_ARG;
    ScState* s = ctx->sc;
    if (!s || !s->active) TEC_FAIL(TEC_ERR_STATE, "tec_sc_exchange: tec_sc_begin not called");
    if (!ctx->comm) TEC_FAIL(TEC_ERR_STATE, "tec_sc_exchange: tec_comm_init not called");
    const int world = ctx->comm_world, rank = ctx->comm_rank;
    if (world > SC_MAX_WORLD) TEC_FAIL(TEC_ERR_LIMIT, "tec_sc_exchange: more ranks than SC_MAX_WORLD");
    TEC_CUDA(cudaSetDevice(ctx->device));
    ncclComm_t comm = (ncclComm_t)ctx->comm;
    ScArena A(ctx->cache);
    u64* d_tab = nullptr;                                   // [world] survivors per rank, then [world][world] send counts
    TEC_CUDA(A.get(&d_tab, (size_t)world * (world + 2)));
    // ---- job-wide position of my first survivor
    TEC_NCCL(g_nccl.AllGather(s->d_n, d_tab, 1, ncclUint64, comm, ctx->stream));
    std::vector<u64> h_n((size_t)world);
    TEC_CUDA(cudaMemcpyAsync(h_n.data(), d_tab, (size_t)world * 8, cudaMemcpyDeviceToHost, ctx->stream));
    TEC_CUDA(cudaStreamSynchronize(ctx->stream));
    int64_t base = 0;
    for (int r = 0; r < rank; ++r) base += (int64_t)h_n[(size_t)r];
    // ---- 32-byte records grouped by owner rank, file order kept inside a group
    std::vector<int64_t> send((size_t)world);
    void* packed = nullptr;
    int rc = tec_sc_partition_dev(ctx, world, base, send.data(), &packed);
    if (rc) return rc;
    u64* d_send = d_tab + world;
    u64* d_all = d_tab + 2 * world;
    std::vector<u64> h_send((size_t)world);
    for (int r = 0; r < world; ++r) h_send[(size_t)r] = (u64)send[(size_t)r];
    TEC_CUDA(cudaMemcpyAsync(d_send, h_send.data(), (size_t)world * 8, cudaMemcpyHostToDevice, ctx->stream));
    TEC_NCCL(g_nccl.AllGather(d_send, d_all, (size_t)world, ncclUint64, comm, ctx->stream));
    std::vector<u64> h_all((size_t)world * world);
    TEC_CUDA(cudaMemcpyAsync(h_all.data(), d_all, h_all.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    TEC_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<int64_t> recv((size_t)world);
    int64_t n_recv = 0;
    for (int r = 0; r < world; ++r) { recv[(size_t)r] = (int64_t)h_all[(size_t)r * world + rank]; n_recv += recv[(size_t)r]; }
    if (n_recv >= (int64_t)0x7FFFFFF0) TEC_FAIL(TEC_ERR_LIMIT, "tec_sc_exchange: more than 2^31 survivors owned by one rank");
    // ---- all-to-all: each rank's file slice precedes the next rank's, every sender sends in file order and the received
    //      parts are laid out by source rank, so the records arrive ascending in position
    char* dst = nullptr;
    TEC_CUDA(A.get(&dst, (size_t)std::max<int64_t>(n_recv, 1) * sizeof(ScRecord)));
    TEC_NCCL(g_nccl.GroupStart());
    int64_t so = 0, ro = 0;
    for (int r = 0; r < world; ++r) {
        const size_t sb = (size_t)send[(size_t)r] * sizeof(ScRecord), rb = (size_t)recv[(size_t)r] * sizeof(ScRecord);
        if (sb) TEC_NCCL(g_nccl.Send((const char*)packed + so, sb, ncclUint8, r, comm, ctx->stream));
        if (rb) TEC_NCCL(g_nccl.Recv(dst + ro, rb, ncclUint8, r, comm, ctx->stream));
        so += (int64_t)sb;
        ro += (int64_t)rb;
    }
    TEC_NCCL(g_nccl.GroupEnd());
    ctx->launches += 2;
    rc = tec_sc_import_packed_dev(ctx, n_recv, dst);         // synchronises the stream: dst may go back to the cache
    if (rc) return rc;
    s->coll = nullptr;
    s->coll_user = nullptr;
    s->rank = rank;
    s->world = world;
    if (n_owned) *n_owned = n_recv;
    return TEC_OK;
}

__global__ void sc_triple_pack_kernel(int64_t n, const int32_t* __restrict__ ensg, const u32* __restrict__ cell, const int64_t* __restrict__ count,
                                      u64* __restrict__ key, int64_t* __restrict__ cnt) {
    SC_LOOP(i, n) { key[i] = ((u64)(u32)ensg[i] << 32) | cell[i]; cnt[i] = count[i]; }
}
__global__ void sc_triple_unpack_kernel(int64_t n, const u64* __restrict__ skey, const u32* __restrict__ perm, const int64_t* __restrict__ cnt,
                                        int32_t* __restrict__ ensg, u32* __restrict__ cell, int64_t* __restrict__ count) {
    SC_LOOP(i, n) { const u64 k = skey[i]; ensg[i] = (int32_t)(k >> 32); cell[i] = (u32)k; count[i] = cnt[perm[i]]; }
}

// After tec_sc_finalize on every rank: the job's triples on every rank, ascending in (ensg, cell).  Cells are disjoint
// between the ranks, so nothing is added: all-gather (grouped send / receive of variable-length parts) and one sort.
extern "C" int tec_sc_allgather_triples(tec_ctx* ctx, int64_t* n_triples) {
    if (!ctx) return TEC_ERR_ARG;
    ScState* s = ctx->sc;
    if (!s || !s->finalized) TEC_FAIL(TEC_ERR_STATE, "tec_sc_allgather_triples: tec_sc_finalize not called");
    if (!ctx->comm) TEC_FAIL(TEC_ERR_STATE, "tec_sc_allgather_triples: tec_comm_init not called");
    const int world = ctx->comm_world, rank = ctx->comm_rank;
    TEC_CUDA(cudaSetDevice(ctx->device));
    ncclComm_t comm = (ncclComm_t)ctx->comm;
    ScArena A(ctx->cache);
    u64* d_cnt = nullptr;
    TEC_CUDA(A.get(&d_cnt, (size_t)world + 1));
    const u64 mine = (u64)s->n_triples;
    TEC_CUDA(cudaMemcpyAsync(d_cnt + world, &mine, 8, cudaMemcpyHostToDevice, ctx->stream));
    TEC_NCCL(g_nccl.AllGather(d_cnt + world, d_cnt, 1, ncclUint64, comm, ctx->stream));
    std::vector<u64> h((size_t)world);
    TEC_CUDA(cudaMemcpyAsync(h.data(), d_cnt, (size_t)world * 8, cudaMemcpyDeviceToHost, ctx->stream));
    TEC_CUDA(cudaStreamSynchronize(ctx->stream));
    int64_t T = 0, my_off = 0;
    for (int r = 0; r < world; ++r) { if (r == rank) my_off = T; T += (int64_t)h[(size_t)r]; }
    if (T >= (int64_t)0xFFFFFFF0 - (1 << 13)) TEC_FAIL(TEC_ERR_LIMIT, "tec_sc_allgather_triples: more than 2^32 triples");
    u64 *ka = nullptr, *kb = nullptr;
    u32 *va = nullptr, *vb = nullptr, *scratch = nullptr;
    int64_t* cnt = nullptr;
    const size_t Tn = (size_t)std::max<int64_t>(T, 1);
    TEC_CUDA(A.get(&ka, Tn)); TEC_CUDA(A.get(&kb, Tn)); TEC_CUDA(A.get(&va, Tn)); TEC_CUDA(A.get(&vb, Tn)); TEC_CUDA(A.get(&cnt, Tn));
    if (s->n_triples) sc_triple_pack_kernel<<<SC_GRID(s->n_triples)>>>(s->n_triples, s->t_ensg, s->t_cell, s->t_count, ka + my_off, cnt + my_off);
    TEC_NCCL(g_nccl.GroupStart());
    int64_t off = 0;
    for (int r = 0; r < world; ++r) {
        const int64_t nr = (int64_t)h[(size_t)r];
        if (r != rank) {
            if (s->n_triples) {
                TEC_NCCL(g_nccl.Send(ka + my_off, (size_t)s->n_triples, ncclUint64, r, comm, ctx->stream));
                TEC_NCCL(g_nccl.Send(cnt + my_off, (size_t)s->n_triples, ncclInt64, r, comm, ctx->stream));
            }
            if (nr) {
                TEC_NCCL(g_nccl.Recv(ka + off, (size_t)nr, ncclUint64, r, comm, ctx->stream));
                TEC_NCCL(g_nccl.Recv(cnt + off, (size_t)nr, ncclInt64, r, comm, ctx->stream));
            }
        }
        off += nr;
    }
    TEC_NCCL(g_nccl.GroupEnd());
    if (T) {
        sc_iota_kernel<<<SC_GRID(T)>>>(T, va);
        const RdxPlan plan = rdx_plan(T, ctx->n_sm);
        TEC_CUDA(A.get(&scratch, plan.counts_bytes / 4));
        // stable LSD: by cell, then by ensg
        const int cell_bits = std::max(1, ceil_log2_i64(std::max<int64_t>(s->n_wl, 2)));
        const int ensg_bits = std::max(1, ceil_log2_i64(std::max<int64_t>(ctx->idx.n_ensg, 2)));
        bool in_b = false, in_b2 = false;
        int p1 = 0, p2 = 0;
        TEC_CUDA((rdx_sort<u64, true>(ka, va, kb, vb, T, 0, cell_bits, ctx->n_sm, scratch, ctx->stream, &in_b, &p1)));
        u64 *k1 = in_b ? kb : ka, *k2 = in_b ? ka : kb;
        u32 *v1 = in_b ? vb : va, *v2 = in_b ? va : vb;
        TEC_CUDA((rdx_sort<u64, true>(k1, v1, k2, v2, T, 32, 32 + ensg_bits, ctx->n_sm, scratch, ctx->stream, &in_b2, &p2)));
        const u64* sk = in_b2 ? k2 : k1;
        const u32* sv = in_b2 ? v2 : v1;
        ctx->launches += 3 + RDX_LAUNCHES_PER_PASS * (p1 + p2);
        int32_t* n_ensg = nullptr; u32* n_cell = nullptr; int64_t* n_count = nullptr;
        TEC_CUDA(ctx->cache.get((void**)&n_ensg, Tn * 4));
        TEC_CUDA(ctx->cache.get((void**)&n_cell, Tn * 4));
        TEC_CUDA(ctx->cache.get((void**)&n_count, Tn * 8));
        sc_triple_unpack_kernel<<<SC_GRID(T)>>>(T, sk, sv, cnt, n_ensg, n_cell, n_count);
        TEC_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->cache.put(s->t_ensg); ctx->cache.put(s->t_cell); ctx->cache.put(s->t_count);
        s->t_ensg = n_ensg; s->t_cell = n_cell; s->t_count = n_count;
    } else {
        TEC_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    s->n_triples = T;
    if (n_triples) *n_triples = T;
    return TEC_OK;
}
