/*
 * oracle/rle.c — TEST INFRASTRUCTURE (see bxw_oracle.h). The VChunk RLE codec and its streaming iterator,
 * restated from bxw_world/src/lib.rs:263-479. Pinned by the reference KATs at lib.rs:486-546.
 */
#include "bxw_oracle.h"

size_t bxo_compress_rle(const uint32_t *data, size_t n, uint32_t *out) {
    /* lib.rs:263-294 */
    size_t o = 0;
    int have_rle = 0;
    uint32_t rle_elem = 0, rle_len = 0;
    int have_prev = 0;
    uint32_t prev = 0;
    for (size_t i = 0; i < n; i++) {
        uint32_t v = data[i];
        if (have_rle) {
            if (rle_elem == v) {
                rle_len += 1;
            } else {
                out[o++] = rle_len;
                out[o++] = v;
                have_prev = 1;
                prev = v;
                have_rle = 0;
            }
        } else {
            out[o++] = v;
            if (have_prev && prev == v) {
                have_prev = 0;
                rle_len = 0;
                have_rle = 1;
                rle_elem = v;
            } else {
                have_prev = 1;
                prev = v;
            }
        }
    }
    if (have_rle) out[o++] = rle_len;
    return o;
}

int bxo_decompress_rle(const uint32_t *data, size_t len, uint32_t out[BXO_CHUNK_DIM3]) {
    /* lib.rs:304-345 */
    if (len < 3) return 4;
    const size_t cap = BXO_CHUNK_DIM3;
    uint32_t first = data[0];
    out[0] = first;
    int have_prev = 1;
    uint32_t prev = first;
    size_t pos = 1;
    size_t i = 1;
    while (i < len) {
        uint32_t e = data[i++];
        if (pos < cap) {
            out[pos] = e;
        } else {
            return 1;
        }
        pos += 1;
        if (have_prev && prev == e) {
            have_prev = 0;
            if (i >= len) return 3;
            size_t extra = data[i++];
            if (pos + extra > cap) return 2;
            for (size_t k = 0; k < extra; k++) out[pos + k] = e;
            pos += extra;
        } else {
            have_prev = 1;
            prev = e;
        }
    }
    return pos == cap ? 0 : 4;
}

void bxo_rle_iter_init(bxo_rle_iter *it, const uint32_t *data, size_t len) {
    /* lib.rs:364-372 (asserts len > 2) */
    it->rem = data;
    it->rem_len = len;
    it->state = 0;
    it->prev = it->what = it->count = 0;
    it->pos = 0;
}

static int advance(bxo_rle_iter *it, uint32_t data, uint32_t *datum, uint32_t *idx) {
    /* lib.rs:405-412 (the BlockPosition by-product is ignored by every caller on the path) */
    *datum = data;
    *idx = it->pos;
    it->pos += 1;
    return 1;
}

int bxo_rle_iter_next(bxo_rle_iter *it, uint32_t *datum, uint32_t *idx) {
    /* lib.rs:418-469 */
    for (;;) {
        if (it->pos >= BXO_CHUNK_DIM3) return 0;
        switch (it->state) {
        case 0: { /* First */
            uint32_t ret = it->rem[0];
            it->state = 1;
            it->prev = ret;
            it->rem += 1;
            it->rem_len -= 1;
            if (it->rem_len == 0) it->state = 3;
            return advance(it, ret, datum, idx);
        }
        case 1: { /* Middle { prev } */
            uint32_t ret = it->rem[0];
            if (it->prev == ret) {
                it->state = 2;
                it->what = ret;
                it->count = it->rem[1];
                it->rem += 2;
                it->rem_len -= 2;
            } else {
                it->prev = ret;
                it->rem += 1;
                it->rem_len -= 1;
                if (it->rem_len == 0) it->state = 3;
            }
            return advance(it, ret, datum, idx);
        }
        case 2: /* InRepetition */
            if (it->count == 0) {
                it->state = 0;
                if (it->rem_len == 0) it->state = 3;
                continue; /* self.next() */
            }
            it->count -= 1;
            return advance(it, it->what, datum, idx);
        default:
            return 0;
        }
    }
}

int bxo_rle_iter_skip_until(bxo_rle_iter *it, uint32_t target, uint32_t *datum, uint32_t *idx) {
    /* lib.rs:374-403 */
    uint32_t d, i;
    while (it->pos < target) {
        if (it->state == 3) return 0;
        if (it->state == 2) {
            uint32_t endpos = it->pos + it->count;
            if (endpos < target) {
                it->count = 0;
                it->pos = endpos;
                bxo_rle_iter_next(it, &d, &i);
            } else {
                uint32_t needed = target - it->pos;
                it->count -= needed;
                it->pos += needed;
            }
        } else {
            bxo_rle_iter_next(it, &d, &i);
        }
    }
    return bxo_rle_iter_next(it, datum, idx);
}
