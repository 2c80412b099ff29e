/* ORACLE — test infrastructure, not product code.
 *
 * Restatement of the APPROXIMATE index the reference queries: chromadb==1.4.1
 * (requirements.txt:33) keeps collection "rag_dpo_chunks" in an hnswlib-style
 * HNSW graph with "hnsw:space": "cosine" (src/processing/create_chromadb_index.py:100-106)
 * and library defaults M = 16, ef_construction = 100, ef_search = 100; collection.query
 * (src/rag/retriever.py:215-220, 380-385) searches it with ef = max(ef_search, n_results).
 * Neither chromadb nor hnswlib is present in /root/reference or installable here, so this
 * follows the published algorithm (Malkov & Yashunin, "Efficient and robust approximate
 * nearest neighbor search using Hierarchical Navigable Small World graphs", Alg. 1-5, with
 * hnswlib's neighbour-selection heuristic and its bidirectional-link pruning).  PARITY
 * UNPINNED: level draws use this file's own generator (hnswlib: std::default_random_engine,
 * seed 100), so the graph is a statistical twin, not a bit copy.  Its only use is to report
 * recall@k of such an index against the exact search (tools/bench_c1.py, tests).
 *
 * Distance = 1 - <a, b> over unit vectors (hnswlib's cosine space after normalisation).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int n, d, M, M0, efc;
    const float* x;     /* n x d, unit rows (borrowed) */
    int* level;         /* n */
    int** links;        /* links[i]: for level l, block at offset off(l): [count, ids...] */
    int entry, max_level;
    uint64_t rng;
    uint32_t* visited;  /* epoch marks */
    uint32_t epoch;
} hnsw_t;

typedef struct { float dist; int id; } cand_t;

/* 16 interleaved partial sums: vectorises at any SIMD width; runtime dispatch keeps the .so portable */
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target_clones("avx512f", "avx2", "default")))
#endif
static float dot_f32(const float* a, const float* b, int d) {
    float s[16] = {0};
    for (int i = 0; i < d; i += 16)
        for (int u = 0; u < 16; ++u) s[u] += a[i + u] * b[i + u];
    float t = 0.f;
    for (int u = 0; u < 16; ++u) t += s[u];
    return t;
}

static float dist_fn(const hnsw_t* h, const float* a, int j) {
    return 1.0f - dot_f32(a, h->x + (size_t)j * h->d, h->d);
}

static int* link_block(const hnsw_t* h, int i, int l) {      /* [count, ids...] of node i at level l */
    return l == 0 ? h->links[i] : h->links[i] + (1 + h->M0) + (size_t)(l - 1) * (1 + h->M);
}

/* ---- binary heaps over cand_t -------------------------------------------------------------- */
typedef struct { cand_t* a; int n, cap, max_heap; } heap_t;
static int heap_before(const heap_t* hp, cand_t p, cand_t q) {   /* p should sit above q */
    return hp->max_heap ? (p.dist > q.dist) : (p.dist < q.dist);
}
static void heap_push(heap_t* hp, cand_t c) {
    if (hp->n == hp->cap) { hp->cap = hp->cap ? 2 * hp->cap : 64; hp->a = (cand_t*)realloc(hp->a, sizeof(cand_t) * hp->cap); }
    int i = hp->n++;
    hp->a[i] = c;
    while (i > 0) {
        int p = (i - 1) / 2;
        if (!heap_before(hp, hp->a[i], hp->a[p])) break;
        cand_t t = hp->a[i]; hp->a[i] = hp->a[p]; hp->a[p] = t;
        i = p;
    }
}
static cand_t heap_pop(heap_t* hp) {
    cand_t top = hp->a[0];
    hp->a[0] = hp->a[--hp->n];
    int i = 0;
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < hp->n && heap_before(hp, hp->a[l], hp->a[m])) m = l;
        if (r < hp->n && heap_before(hp, hp->a[r], hp->a[m])) m = r;
        if (m == i) break;
        cand_t t = hp->a[i]; hp->a[i] = hp->a[m]; hp->a[m] = t;
        i = m;
    }
    return top;
}

/* Alg. 2: ef closest to q found from entry points ep at one level; result in `top` (max-heap, <= ef) */
static void search_layer(hnsw_t* h, const float* q, cand_t ep, int ef, int level, heap_t* top) {
    heap_t cand = {0, 0, 0, 0};
    top->n = 0;
    top->max_heap = 1;
    if (++h->epoch == 0) { memset(h->visited, 0, sizeof(uint32_t) * h->n); h->epoch = 1; }
    h->visited[ep.id] = h->epoch;
    heap_push(&cand, ep);
    heap_push(top, ep);
    while (cand.n > 0) {
        cand_t c = heap_pop(&cand);
        if (top->n >= ef && c.dist > top->a[0].dist) break;
        const int* blk = link_block(h, c.id, level);
        for (int e = 1; e <= blk[0]; ++e) {
            int nb = blk[e];
            if (h->visited[nb] == h->epoch) continue;
            h->visited[nb] = h->epoch;
            cand_t nc = {dist_fn(h, q, nb), nb};
            if (top->n < ef || nc.dist < top->a[0].dist) {
                heap_push(&cand, nc);
                heap_push(top, nc);
                if (top->n > ef) heap_pop(top);
            }
        }
    }
    free(cand.a);
}

static int cmp_cand(const void* a, const void* b) {
    const cand_t* p = (const cand_t*)a;
    const cand_t* q = (const cand_t*)b;
    if (p->dist != q->dist) return p->dist < q->dist ? -1 : 1;
    return p->id - q->id;
}

/* hnswlib's heuristic (Alg. 4 without extension): walk candidates by increasing distance, keep one unless it
 * is closer to an already kept neighbour than to the base point */
static int select_neighbors(const hnsw_t* h, cand_t* c, int nc, int M, int* out) {
    qsort(c, nc, sizeof(cand_t), cmp_cand);
    int kept = 0;
    for (int i = 0; i < nc && kept < M; ++i) {
        int good = 1;
        const float* ci = h->x + (size_t)c[i].id * h->d;
        for (int j = 0; j < kept; ++j)
            if (dist_fn(h, ci, out[j]) < c[i].dist) { good = 0; break; }
        if (good) out[kept++] = c[i].id;
    }
    return kept;
}

static double rnd01(hnsw_t* h) {       /* xorshift64*: (0,1) */
    h->rng ^= h->rng >> 12; h->rng ^= h->rng << 25; h->rng ^= h->rng >> 27;
    return ((h->rng * 2685821657736338717ULL) >> 11) * (1.0 / 9007199254740992.0) + 1e-18;
}

static void insert(hnsw_t* h, int i, heap_t* top, cand_t* scratch, int* sel) {
    const float* q = h->x + (size_t)i * h->d;
    const int lvl = h->level[i];
    if (h->entry < 0) { h->entry = i; h->max_level = lvl; return; }
    cand_t ep = {dist_fn(h, q, h->entry), h->entry};
    for (int l = h->max_level; l > lvl; --l) {          /* greedy descent, ef = 1 */
        int changed = 1;
        while (changed) {
            changed = 0;
            const int* blk = link_block(h, ep.id, l);
            for (int e = 1; e <= blk[0]; ++e) {
                float dd = dist_fn(h, q, blk[e]);
                if (dd < ep.dist) { ep.dist = dd; ep.id = blk[e]; changed = 1; }
            }
        }
    }
    for (int l = lvl < h->max_level ? lvl : h->max_level; l >= 0; --l) {
        search_layer(h, q, ep, h->efc, l, top);
        int nc = top->n;
        memcpy(scratch, top->a, sizeof(cand_t) * nc);
        const int Mmax = l == 0 ? h->M0 : h->M;
        int ns = select_neighbors(h, scratch, nc, h->M, sel);
        int* mine = link_block(h, i, l);
        mine[0] = ns;
        memcpy(mine + 1, sel, sizeof(int) * ns);
        ep = scratch[0];                                   /* closest found: entry of the next level */
        for (int s = 0; s < ns; ++s) {                     /* back links, pruned by the same heuristic */
            int nb = sel[s];
            int* blk = link_block(h, nb, l);
            if (blk[0] < Mmax) { blk[++blk[0]] = i; continue; }
            const float* xb = h->x + (size_t)nb * h->d;
            cand_t* cc = scratch + nc;                     /* scratch holds efc + Mmax + 1 entries */
            int m = 0;
            cc[m].dist = dist_fn(h, xb, i); cc[m++].id = i;
            for (int e = 1; e <= blk[0]; ++e) { cc[m].dist = dist_fn(h, xb, blk[e]); cc[m++].id = blk[e]; }
            int tmp[512];
            int kept = select_neighbors(h, cc, m, Mmax, tmp);
            blk[0] = kept;
            memcpy(blk + 1, tmp, sizeof(int) * kept);
        }
    }
    if (lvl > h->max_level) { h->max_level = lvl; h->entry = i; }
}

void* hnsw_build(const float* x, int n, int d, int M, int ef_construction, uint64_t seed) {
    hnsw_t* h = (hnsw_t*)calloc(1, sizeof(hnsw_t));
    h->n = n; h->d = d; h->M = M; h->M0 = 2 * M; h->efc = ef_construction; h->x = x;
    h->entry = -1; h->max_level = -1; h->rng = seed * 0x9E3779B97F4A7C15ULL + 0x1234567ULL;
    h->level = (int*)malloc(sizeof(int) * n);
    h->links = (int**)malloc(sizeof(int*) * n);
    h->visited = (uint32_t*)calloc(n, sizeof(uint32_t));
    const double mult = 1.0 / log((double)M);
    for (int i = 0; i < n; ++i) {
        h->level[i] = (int)(-log(rnd01(h)) * mult);
        h->links[i] = (int*)calloc((size_t)(1 + h->M0) + (size_t)h->level[i] * (1 + M), sizeof(int));
    }
    heap_t top = {0, 0, 0, 1};
    /* efc candidates of a layer search + (Mmax + 1) entries while a neighbour's links are pruned */
    cand_t* scratch = (cand_t*)malloc(sizeof(cand_t) * (size_t)(2 * ef_construction + 2 * h->M0 + 16));
    int* sel = (int*)malloc(sizeof(int) * (size_t)(h->M0 + 1));
    for (int i = 0; i < n; ++i) insert(h, i, &top, scratch, sel);
    free(top.a); free(scratch); free(sel);
    return h;
}

/* k nearest by the graph (ef = max(ef_search, k)); ids ascending by distance, returns how many */
int hnsw_search(void* hv, const float* q, int k, int ef_search, int* out_ids, float* out_dist) {
    hnsw_t* h = (hnsw_t*)hv;
    if (h->entry < 0) return 0;
    cand_t ep = {dist_fn(h, q, h->entry), h->entry};
    for (int l = h->max_level; l > 0; --l) {
        int changed = 1;
        while (changed) {
            changed = 0;
            const int* blk = link_block(h, ep.id, l);
            for (int e = 1; e <= blk[0]; ++e) {
                float dd = dist_fn(h, q, blk[e]);
                if (dd < ep.dist) { ep.dist = dd; ep.id = blk[e]; changed = 1; }
            }
        }
    }
    heap_t top = {0, 0, 0, 1};
    const int ef = ef_search > k ? ef_search : k;
    search_layer(h, q, ep, ef, 0, &top);
    qsort(top.a, top.n, sizeof(cand_t), cmp_cand);
    const int m = top.n < k ? top.n : k;
    for (int i = 0; i < m; ++i) { out_ids[i] = top.a[i].id; out_dist[i] = top.a[i].dist; }
    free(top.a);
    return m;
}

void hnsw_free(void* hv) {
    hnsw_t* h = (hnsw_t*)hv;
    for (int i = 0; i < h->n; ++i) free(h->links[i]);
    free(h->links); free(h->level); free(h->visited); free(h);
}
