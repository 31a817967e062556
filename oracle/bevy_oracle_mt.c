/*
 * bevy_oracle_mt.c -- multithreaded CPU BASELINE (TEST/BENCH INFRASTRUCTURE).
 *
 * The same restatement as bevy_oracle.c (included verbatim below), parallelised
 * the way the reference parallelises it, so bench.py can time "the reference's
 * algorithm on the box's host cores" (option 2):
 *   - propagate: parallel over roots, like root_query.par_iter_mut() + the work
 *     sharing workers (crates/bevy_transform/src/systems.rs:522-581); flat
 *     entities in parallel like sync_simple_transforms (systems.rs:56-63)
 *   - cull: per view, parallel over contiguous row ranges with thread-local
 *     queues, then a serial merge + sort_unstable
 *     (crates/bevy_camera/src/visibility/mod.rs:786-874; batching as in
 *     crates/bevy_ecs/src/query/state.rs:1661-1676)
 *   - cluster: single threaded in the reference (assign.rs:137) -> orc_assign_lights_to_clusters
 * Results are bit-identical to the serial oracle (tests/test_oracle_mt.py).
 * Built -O3 -march=native -ffp-contract=off -fopenmp.
 */
#include "bevy_oracle.c"
#include <omp.h>

ORC_API int orc_mt_threads(void) { return omp_get_max_threads(); }
ORC_API void orc_mt_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

/* Persistent per-topology state, like the reference's Children components and Local<> scratch buffers: the children
 * lists and the scratch arrays are built once per hierarchy (orc_mt_prepare), not once per frame. */
static uint32_t g_n = 0, *g_first = NULL, *g_kids = NULL;
static uint8_t *g_dirty = NULL;
static sort_item *g_items = NULL;
static const uint32_t *g_parent = NULL;

ORC_API int orc_mt_prepare(uint32_t n, const uint32_t *parent) {
    free(g_first); free(g_kids); free(g_dirty); free(g_items);
    g_first = (uint32_t *)calloc((size_t)n + 1, sizeof(uint32_t));
    g_kids = (uint32_t *)malloc((size_t)(n ? n : 1) * sizeof(uint32_t));
    g_dirty = (uint8_t *)calloc(n ? n : 1, 1);
    g_items = (sort_item *)malloc((size_t)(n ? n : 1) * sizeof(sort_item));
    g_n = n; g_parent = parent;
    for (uint32_t r = 0; r < n; ++r) {
        uint32_t p = parent[r];
        if (p < n) g_first[p + 1]++;
        else if (p != ORC_NO_PARENT && p != ORC_DETACHED) return -1;
    }
    for (uint32_t r = 0; r < n; ++r) g_first[r + 1] += g_first[r];
    uint32_t *cur = (uint32_t *)malloc((size_t)(n ? n : 1) * sizeof(uint32_t));
    memcpy(cur, g_first, (size_t)n * sizeof(uint32_t));
    for (uint32_t r = 0; r < n; ++r) { uint32_t p = parent[r]; if (p < n) g_kids[cur[p]++] = r; }
    free(cur);
    /* first touch of the scratch arrays in parallel */
    #pragma omp parallel for schedule(static)
    for (uint32_t r = 0; r < n; ++r) { g_items[r].key = 0; g_items[r].row = 0; }
    return 0;
}

ORC_API int orc_propagate_mt(uint32_t n, const uint32_t *parent, const float *trs, float *gt,
                             const uint8_t *tchanged, const uint8_t *gt_ext_changed, int static_opt,
                             uint8_t *changed) {
    if (n == 0) return 0;
    if (g_n != n || g_parent != parent) { int rc = orc_mt_prepare(n, parent); if (rc) return rc; }
    uint32_t *first = g_first, *kids = g_kids;
    uint8_t *dirty = g_dirty;
    #pragma omp parallel for schedule(static)
    for (uint32_t r = 0; r < n; ++r) { changed[r] = 0; dirty[r] = 0; }
    if (static_opt) {
        /* mark_dirty_trees: benign-race ancestor marking (the reference uses fetch_or, systems.rs:208-223) */
        #pragma omp parallel for schedule(static)
        for (uint32_t r = 0; r < n; ++r) {
            if (!tchanged[r]) continue;
            uint32_t c = r;
            while (!__atomic_exchange_n(&dirty[c], 1, __ATOMIC_RELAXED)) {
                uint32_t p = parent[c];
                if (p >= n) break;
                c = p;
            }
        }
    }
    #pragma omp parallel
    {
        uint32_t cap = 1024, *stack = (uint32_t *)malloc(cap * sizeof(uint32_t));
        #pragma omp for schedule(dynamic, 64)
        for (uint32_t r = 0; r < n; ++r) {
            if (parent[r] != ORC_NO_PARENT) continue;
            int has_children = first[r + 1] > first[r];
            if (!has_children) {
                if (tchanged[r]) {
                    aff a = aff_from_trs(trs + (size_t)r * 10);
                    aff_store(gt + (size_t)r * 12, &a);
                    changed[r] = 1;
                }
                continue;
            }
            if (static_opt && !dirty[r]) continue;
            {
                aff a = aff_from_trs(trs + (size_t)r * 10);
                aff_store(gt + (size_t)r * 12, &a);
                changed[r] = 1;
            }
            uint32_t sp = 0;
            stack[sp++] = r;
            while (sp) {
                uint32_t p = stack[--sp];
                int p_changed = changed[p] || (gt_ext_changed && gt_ext_changed[p]);
                aff pg = aff_load(gt + (size_t)p * 12);
                for (uint32_t k = first[p]; k < first[p + 1]; ++k) {
                    uint32_t c = kids[k];
                    if (static_opt && !dirty[c] && !p_changed) continue;
                    aff l = aff_from_trs(trs + (size_t)c * 10);
                    aff g = aff_mul(&pg, &l);
                    float tmp[12];
                    aff_store(tmp, &g);
                    if (gt_neq(tmp, gt + (size_t)c * 12)) {
                        memcpy(gt + (size_t)c * 12, tmp, sizeof tmp);
                        changed[c] = 1;
                    }
                    if (first[c + 1] > first[c]) {
                        if (sp == cap) { cap *= 2; stack = (uint32_t *)realloc(stack, cap * sizeof(uint32_t)); }
                        stack[sp++] = c;
                    }
                }
            }
        }
        free(stack);
    }
    return 0;
}

ORC_API int orc_cull_mt(uint32_t n, const float *gt, const float *bounds, const uint8_t *flags,
                        const uint64_t *layer_mask, const uint32_t *range_mask, const uint8_t *class_mask,
                        const uint64_t *entity_bits, uint8_t *vv, uint8_t *vv_changed,
                        uint32_t n_views, const float *view_planes, const uint64_t *view_layers,
                        const uint8_t *view_flags, const int8_t *view_range_index,
                        uint32_t *visible_rows, uint32_t *visible_count) {
    #pragma omp parallel for schedule(static)
    for (uint32_t r = 0; r < n; ++r) {
        if (!(flags[r] & F_NO_CPU_CULLING)) vv[r] = (uint8_t)((vv[r] & 1u) << 1);
        vv_changed[r] = 0;
    }
    int nt = omp_get_max_threads();
    sort_item *items = (g_n == n && g_items) ? g_items : (sort_item *)malloc((size_t)(n ? n : 1) * sizeof(sort_item));
    uint32_t *tcount = (uint32_t *)calloc((size_t)nt + 1, sizeof(uint32_t));
    for (uint32_t v = 0; v < n_views; ++v) {
        if (!(view_flags[v] & VIEW_ACTIVE)) { visible_count[v] = 0xFFFFFFFFu; continue; }
        v4 hs[6]; memcpy(hs, view_planes + (size_t)v * 24, sizeof hs);
        memset(tcount, 0, ((size_t)nt + 1) * sizeof(uint32_t));
        /* contiguous row ranges per thread; each thread appends into its own slice of `items`
         * (its range start), which plays the role of the thread-local queue */
        #pragma omp parallel num_threads(nt)
        {
            int t = omp_get_thread_num();
            uint32_t lo = (uint32_t)((uint64_t)n * t / nt), hi = (uint32_t)((uint64_t)n * (t + 1) / nt);
            uint32_t cnt = 0;
            for (uint32_t r = lo; r < hi; ++r) {
                uint8_t f = flags[r];
                if (f & F_NO_CPU_CULLING) continue;
                uint64_t el = layer_mask ? layer_mask[r] : 1ull;
                if (!entity_visible_in_view(r, gt, bounds, f, el, view_layers[v], range_mask,
                                            view_range_index ? view_range_index[v] : -1, hs,
                                            (view_flags[v] & VIEW_NO_CPU_CULLING) != 0))
                    continue;
                if (!(vv[r] & 1u)) {
                    if (!(vv[r] & 2u)) vv_changed[r] = 1;
                    vv[r] |= 1u;
                }
                if (class_mask[r]) { items[lo + cnt].key = entity_bits[r]; items[lo + cnt].row = r; cnt++; }
            }
            tcount[t + 1] = cnt;
        }
        /* serial drain + sort_unstable (visibility/mod.rs:861-874) */
        uint32_t total = 0;
        for (int t = 0; t < nt; ++t) {
            uint32_t lo = (uint32_t)((uint64_t)n * t / nt);
            if (total != lo) memmove(items + total, items + lo, (size_t)tcount[t + 1] * sizeof(sort_item));
            total += tcount[t + 1];
        }
        qsort(items, total, sizeof(sort_item), cmp_sort_item);
        for (uint32_t i = 0; i < total; ++i) visible_rows[(size_t)v * n + i] = items[i].row;
        visible_count[v] = total;
    }
    #pragma omp parallel for schedule(static)
    for (uint32_t r = 0; r < n; ++r) {
        if (flags[r] & F_NO_CPU_CULLING) continue;
        if ((vv[r] & 3u) == 2u) { vv[r] = 0; vv_changed[r] = 1; }
    }
    if (items != g_items) free(items);
    free(tcount);
    return 0;
}
