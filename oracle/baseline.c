/*
 * oracle/baseline.c — TEST INFRASTRUCTURE (see bxw_oracle.h). Drives the oracle the way the reference drives its
 * CPU path, for the reported CPU baseline only:
 *   - one task per chunk: StdGenerator::generate_chunk -> VChunk::compress (generation.rs:96-148)
 *   - one task per chunk: gather 27 neighbours, is_chunk_trivial early-out, mesh_from_chunk reading the
 *     neighbours through RleVoxelIterator (voxrender.rs:224-279, voxmesh.rs:45-190)
 *   - a shared-queue pool of worker threads (bxw_util/src/taskpool.rs:54-169); each worker owns a thread-local
 *     CellGen (stdgen.rs:285-288, 328-331).
 */
#include "bxw_oracle.h"

#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct {
    uint32_t *words;
    size_t len;
} rle_chunk;

typedef struct {
    uint64_t seed;
    int32_t cx0, cy0, cz0; /* origin of the shell-padded box */
    int sx, sy, sz;        /* shell-padded dims */
    rle_chunk *store;      /* sx*sy*sz, index x + sx*(z + sz*y) */
    bxo_registry reg;
    uint16_t ids[5];
    atomic_long next;
    long total;
    int phase; /* 0 gen, 1 mesh */
    atomic_ullong tv, ti, nontrivial;
} job;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void *worker(void *arg) {
    job *j = (job *)arg;
    if (j->phase == 0) {
        bxo_stdgen *gen = bxo_stdgen_new(j->seed);
        uint32_t *dense = (uint32_t *)malloc(sizeof(uint32_t) * BXO_CHUNK_DIM3);
        uint32_t *tmp = (uint32_t *)malloc(sizeof(uint32_t) * BXO_CHUNK_DIM3 * 2);
        for (;;) {
            long t = atomic_fetch_add(&j->next, 1);
            if (t >= j->total) break;
            int x = (int)(t % j->sx), z = (int)((t / j->sx) % j->sz), y = (int)(t / ((long)j->sx * j->sz));
            bxo_generate_chunk(gen, j->cx0 + x, j->cy0 + y, j->cz0 + z, j->ids, dense);
            size_t n = bxo_compress_rle(dense, BXO_CHUNK_DIM3, tmp);
            uint32_t *w = (uint32_t *)malloc(n * sizeof(uint32_t));
            memcpy(w, tmp, n * sizeof(uint32_t));
            j->store[t].words = w;
            j->store[t].len = n;
        }
        free(dense);
        free(tmp);
        bxo_stdgen_free(gen);
    } else {
        int nx = j->sx - 2, nz = j->sz - 2;
        for (;;) {
            long t = atomic_fetch_add(&j->next, 1);
            if (t >= j->total) break;
            int x = 1 + (int)(t % nx), z = 1 + (int)((t / nx) % nz), y = 1 + (int)(t / ((long)nx * nz));
            const uint32_t *rle[27];
            size_t len[27];
            int k = 0;
            for (int dy = -1; dy <= 1; dy++)
                for (int dz = -1; dz <= 1; dz++)
                    for (int dx = -1; dx <= 1; dx++) {
                        long idx = (x + dx) + (long)j->sx * ((z + dz) + (long)j->sz * (y + dy));
                        rle[k] = j->store[idx].words;
                        len[k] = j->store[idx].len;
                        k++;
                    }
            if (bxo_is_chunk_trivial(&j->reg, rle[13], len[13])) continue;
            bxo_mesh m;
            if (bxo_mesh_from_chunk(&j->reg, rle, len, &m) == 0) {
                atomic_fetch_add(&j->tv, m.n_vertices);
                atomic_fetch_add(&j->ti, m.n_indices);
                atomic_fetch_add(&j->nontrivial, 1);
                bxo_mesh_free(&m);
            }
        }
    }
    return 0;
}

static void run_phase(job *j, int threads) {
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    atomic_store(&j->next, 0);
    for (int i = 0; i < threads; i++) pthread_create(&th[i], 0, worker, j);
    for (int i = 0; i < threads; i++) pthread_join(th[i], 0);
    free(th);
}

int bxo_baseline_region(uint64_t seed, int32_t cx0, int32_t cy0, int32_t cz0, int nx, int ny, int nz, int threads,
                        bxo_baseline_result *out) {
    if (nx <= 0 || ny <= 0 || nz <= 0 || threads <= 0) return -1;
    job j;
    memset(&j, 0, sizeof j);
    bxo_block_def defs[8];
    bxo_standard_registry(defs);
    bxo_shape_table(0);
    j.reg.n = 8;
    j.reg.defs = defs;
    j.ids[0] = 0; /* void */
    j.ids[1] = 1; /* grass */
    j.ids[2] = 3; /* dirt */
    j.ids[3] = 4; /* stone */
    j.ids[4] = 2; /* snow_grass */
    j.seed = seed;
    j.cx0 = cx0 - 1;
    j.cy0 = cy0 - 1;
    j.cz0 = cz0 - 1;
    j.sx = nx + 2;
    j.sy = ny + 2;
    j.sz = nz + 2;
    long nstore = (long)j.sx * j.sy * j.sz;
    j.store = (rle_chunk *)calloc((size_t)nstore, sizeof(rle_chunk));
    if (!j.store) return -2;

    j.phase = 0;
    j.total = nstore;
    double t0 = now_s();
    run_phase(&j, threads);
    double t1 = now_s();
    j.phase = 1;
    j.total = (long)nx * ny * nz;
    run_phase(&j, threads);
    double t2 = now_s();

    out->gen_seconds = t1 - t0;
    out->mesh_seconds = t2 - t1;
    out->chunks = (uint64_t)j.total;
    out->gen_chunks = (uint64_t)nstore;
    out->total_vertices = atomic_load(&j.tv);
    out->total_indices = atomic_load(&j.ti);
    out->nontrivial = atomic_load(&j.nontrivial);
    out->threads = threads;
    for (long i = 0; i < nstore; i++) free(j.store[i].words);
    free(j.store);
    return 0;
}
