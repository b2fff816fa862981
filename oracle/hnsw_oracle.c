/* ORACLE - test infrastructure only.  Never linked into, imported by or executed from the product.
 *
 * CPU restatement of the APPROXIMATE index the reference actually queries: hnswlib 0.8.0
 * `Index(space='cosine', dim)` with `init_index(max_elements=100000, ef_construction=200, M=16)`,
 * `set_ef(200)` (modules/hnsw_manager.py:20,29-30; one `add_items` per row, :127,137; `knn_query`, :147,237).
 * hnswlib is not vendored under /root/reference and its wheel is not installable here, so this follows the
 * published algorithm (Malkov & Yashunin, "Efficient and robust approximate nearest neighbor search using
 * Hierarchical Navigable Small World graphs") with hnswlib's choices, restated from its hnswalg.h:
 *   - level of a new element = floor(-ln(U(0,1)) * 1/ln(M)), U from std::default_random_engine(seed 100)
 *     (minstd_rand0) through uniform_real_distribution<double>;
 *   - maxM = M, maxM0 = 2 M; insertion = greedy descent with ef 1 down to level+1, then per level
 *     searchBaseLayer(ef_construction) -> getNeighborsByHeuristic2 -> mutual connection, pruning a full
 *     neighbour list with the same heuristic;
 *   - query = greedy descent to level 0, searchBaseLayerST with max(ef, k), k nearest returned ascending;
 *   - cosine space = inner product on rows normalised with 1/(||x|| + 1e-30), distance 1 - <a, b>.
 * PARITY UNPINNED against hnswlib itself (no golden vectors exist; graph construction order and tie handling
 * may differ in detail).  It is used only to REPORT recall@10 of "the reference's HNSW index" against the exact
 * result (tools/recall_report.py, north star), never to gate the product.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int dim, M, maxM, maxM0, efc, max_elements, count, enter, maxlevel;
  double mult;
  uint32_t rng;          /* minstd_rand0 state */
  float* data;           /* [max_elements][dim], normalised */
  int* level;            /* [max_elements] */
  int* link0;            /* [max_elements][1 + maxM0] : count, ids */
  int** linkU;           /* [max_elements] -> [level][1 + maxM] or NULL */
  uint32_t* visited;     /* [max_elements] epoch marks */
  uint32_t epoch;
} hnsw_t;

static float dist_ip(const float* a, const float* b, int d) {
  float s = 0.f;
  for (int i = 0; i < d; ++i) s += a[i] * b[i];
  return 1.0f - s;
}

/* minstd_rand0 + std::generate_canonical<double, 53> (two draws of a 31-bit engine) */
static double next_uniform(hnsw_t* h) {
  const double R = 2147483646.0;
  double sum = 0.0, mul = 1.0;
  for (int k = 0; k < 2; ++k) {
    h->rng = (uint32_t)(((uint64_t)h->rng * 16807ull) % 2147483647ull);
    sum += (double)(h->rng - 1u) * mul;
    mul *= R;
  }
  double r = sum / mul;
  if (r >= 1.0) r = nextafter(1.0, 0.0);
  return r;
}

/* ---- binary heaps of (dist, id) ---------------------------------------------------------------- */
typedef struct { float d; int id; } cand_t;
typedef struct { cand_t* a; int n, cap; } heap_t;
static void heap_init(heap_t* h, int cap) { h->a = (cand_t*)malloc(sizeof(cand_t) * (size_t)cap); h->n = 0; h->cap = cap; }
static void heap_free(heap_t* h) { free(h->a); }
static void heap_grow(heap_t* h) { if (h->n == h->cap) { h->cap *= 2; h->a = (cand_t*)realloc(h->a, sizeof(cand_t) * (size_t)h->cap); } }
/* max_heap != 0: largest distance on top; else smallest on top */
static void heap_push(heap_t* h, cand_t c, int max_heap) {
  heap_grow(h);
  int i = h->n++;
  while (i > 0) {
    int p = (i - 1) / 2;
    int up = max_heap ? (h->a[p].d < c.d) : (h->a[p].d > c.d);
    if (!up) break;
    h->a[i] = h->a[p];
    i = p;
  }
  h->a[i] = c;
}
static cand_t heap_pop(heap_t* h, int max_heap) {
  cand_t top = h->a[0], last = h->a[--h->n];
  int i = 0;
  for (;;) {
    int l = 2 * i + 1, r = l + 1, b = i;
    cand_t bv = last;
    if (l < h->n && (max_heap ? (h->a[l].d > bv.d) : (h->a[l].d < bv.d))) { b = l; bv = h->a[l]; }
    if (r < h->n && (max_heap ? (h->a[r].d > bv.d) : (h->a[r].d < bv.d))) { b = r; bv = h->a[r]; }
    if (b == i) break;
    h->a[i] = h->a[b];
    i = b;
  }
  if (h->n > 0) h->a[i] = last;
  return top;
}

static int* links_of(hnsw_t* h, int id, int lev) { return lev == 0 ? h->link0 + (size_t)id * (1 + h->maxM0) : h->linkU[id] + (size_t)(lev - 1) * (1 + h->maxM); }

/* searchBaseLayer: best-first search on one level; result = max-heap `top` with at most ef entries */
static void search_layer(hnsw_t* h, const float* q, int ep, int ef, int lev, heap_t* top) {
  heap_t cand;
  heap_init(&cand, 256);
  if (++h->epoch == 0) { memset(h->visited, 0, sizeof(uint32_t) * (size_t)h->max_elements); h->epoch = 1; }
  float d = dist_ip(q, h->data + (size_t)ep * h->dim, h->dim);
  cand_t c = {d, ep};
  heap_push(top, c, 1);
  heap_push(&cand, c, 0);
  h->visited[ep] = h->epoch;
  float lower = d;
  while (cand.n > 0) {
    cand_t cur = cand.a[0];
    if (cur.d > lower && top->n == ef) break;
    heap_pop(&cand, 0);
    const int* ll = links_of(h, cur.id, lev);
    for (int j = 1; j <= ll[0]; ++j) {
      const int nb = ll[j];
      if (h->visited[nb] == h->epoch) continue;
      h->visited[nb] = h->epoch;
      const float dn = dist_ip(q, h->data + (size_t)nb * h->dim, h->dim);
      if (top->n < ef || dn < lower) {
        cand_t e = {dn, nb};
        heap_push(&cand, e, 0);
        heap_push(top, e, 1);
        if (top->n > ef) heap_pop(top, 1);
        lower = top->a[0].d;
      }
    }
  }
  heap_free(&cand);
}

/* getNeighborsByHeuristic2: keep a candidate only if it is closer to the query than to every kept one */
static int select_heuristic(hnsw_t* h, cand_t* c, int n, int M, int* out) {
  /* c sorted ascending by distance */
  int kept = 0;
  for (int i = 0; i < n && kept < M; ++i) {
    int good = 1;
    for (int j = 0; j < kept; ++j) {
      const float dd = dist_ip(h->data + (size_t)out[j] * h->dim, h->data + (size_t)c[i].id * h->dim, h->dim);
      if (dd < c[i].d) { good = 0; break; }
    }
    if (good) out[kept++] = c[i].id;
  }
  return kept;
}
static int cmp_cand(const void* a, const void* b) {
  const cand_t *x = (const cand_t*)a, *y = (const cand_t*)b;
  return x->d < y->d ? -1 : (x->d > y->d ? 1 : (x->id - y->id));
}

hnsw_t* fire_hnsw_create(int dim, int max_elements, int M, int ef_construction, unsigned seed) {
  hnsw_t* h = (hnsw_t*)calloc(1, sizeof(hnsw_t));
  h->dim = dim; h->M = M; h->maxM = M; h->maxM0 = 2 * M; h->efc = ef_construction > M ? ef_construction : M;
  h->max_elements = max_elements; h->enter = -1; h->maxlevel = -1;
  h->mult = 1.0 / log((double)M);
  h->rng = seed % 2147483647u; if (h->rng == 0) h->rng = 1;
  h->data = (float*)malloc(sizeof(float) * (size_t)max_elements * dim);
  h->level = (int*)calloc((size_t)max_elements, sizeof(int));
  h->link0 = (int*)calloc((size_t)max_elements * (1 + h->maxM0), sizeof(int));
  h->linkU = (int**)calloc((size_t)max_elements, sizeof(int*));
  h->visited = (uint32_t*)calloc((size_t)max_elements, sizeof(uint32_t));
  return h;
}
void fire_hnsw_destroy(hnsw_t* h) {
  if (!h) return;
  for (int i = 0; i < h->count; ++i) free(h->linkU[i]);
  free(h->data); free(h->level); free(h->link0); free(h->linkU); free(h->visited); free(h);
}
int fire_hnsw_count(const hnsw_t* h) { return h->count; }

static void connect(hnsw_t* h, int id, int lev, int* sel, int nsel) {
  const int maxm = lev == 0 ? h->maxM0 : h->maxM;
  int* mine = links_of(h, id, lev);
  mine[0] = nsel;
  for (int j = 0; j < nsel; ++j) mine[1 + j] = sel[j];
  for (int j = 0; j < nsel; ++j) {
    int* ll = links_of(h, sel[j], lev);
    if (ll[0] < maxm) { ll[1 + ll[0]] = id; ll[0]++; continue; }
    /* full: re-select among the old neighbours + the new element */
    cand_t* c = (cand_t*)malloc(sizeof(cand_t) * (size_t)(maxm + 1));
    const float* base = h->data + (size_t)sel[j] * h->dim;
    for (int t = 0; t < maxm; ++t) { c[t].id = ll[1 + t]; c[t].d = dist_ip(base, h->data + (size_t)ll[1 + t] * h->dim, h->dim); }
    c[maxm].id = id; c[maxm].d = dist_ip(base, h->data + (size_t)id * h->dim, h->dim);
    qsort(c, (size_t)maxm + 1, sizeof(cand_t), cmp_cand);
    int* out = (int*)malloc(sizeof(int) * (size_t)maxm);
    const int k = select_heuristic(h, c, maxm + 1, maxm, out);
    ll[0] = k;
    for (int t = 0; t < k; ++t) ll[1 + t] = out[t];
    free(out); free(c);
  }
}

/* add one row (normalised here like hnswlib's cosine space); label = running count */
int fire_hnsw_add(hnsw_t* h, const float* row) {
  if (h->count >= h->max_elements) return -1;
  const int id = h->count;
  float* v = h->data + (size_t)id * h->dim;
  float s = 0.f;
  for (int i = 0; i < h->dim; ++i) s += row[i] * row[i];
  const float inv = 1.0f / (sqrtf(s) + 1e-30f);
  for (int i = 0; i < h->dim; ++i) v[i] = row[i] * inv;
  const int lev = (int)(-log(next_uniform(h)) * h->mult);
  h->level[id] = lev;
  if (lev > 0) h->linkU[id] = (int*)calloc((size_t)lev * (1 + h->maxM), sizeof(int));
  h->count++;
  if (h->enter < 0) { h->enter = id; h->maxlevel = lev; return id; }
  int cur = h->enter;
  float curd = dist_ip(v, h->data + (size_t)cur * h->dim, h->dim);
  for (int l = h->maxlevel; l > lev; --l) {            /* greedy descent, ef = 1 */
    int changed = 1;
    while (changed) {
      changed = 0;
      const int* ll = links_of(h, cur, l);
      for (int j = 1; j <= ll[0]; ++j) {
        const float d = dist_ip(v, h->data + (size_t)ll[j] * h->dim, h->dim);
        if (d < curd) { curd = d; cur = ll[j]; changed = 1; }
      }
    }
  }
  for (int l = lev < h->maxlevel ? lev : h->maxlevel; l >= 0; --l) {
    heap_t top;
    heap_init(&top, h->efc + 2);
    search_layer(h, v, cur, h->efc, l, &top);
    const int n = top.n;
    cand_t* c = (cand_t*)malloc(sizeof(cand_t) * (size_t)n);
    for (int i = n - 1; i >= 0; --i) c[i] = heap_pop(&top, 1);      /* ascending */
    int m = 0;
    for (int i = 0; i < n; ++i) if (c[i].id != id) c[m++] = c[i];
    int* sel = (int*)malloc(sizeof(int) * (size_t)h->M);
    const int k = select_heuristic(h, c, m, h->M, sel);
    connect(h, id, l, sel, k);
    if (m > 0) cur = c[0].id;
    free(sel); free(c); heap_free(&top);
  }
  if (lev > h->maxlevel) { h->maxlevel = lev; h->enter = id; }
  return id;
}

/* knn_query for nq rows: labels [nq][k] (int64), distances [nq][k] ascending.  Returns 0, or -1 if fewer than k found. */
int fire_hnsw_query(hnsw_t* h, const float* queries, int nq, int k, int ef, long long* labels, float* dists) {
  if (h->count == 0 || k > h->count) return -1;
  float* qn = (float*)malloc(sizeof(float) * (size_t)h->dim);
  const int efs = ef > k ? ef : k;
  int rc = 0;
  for (int qi = 0; qi < nq; ++qi) {
    const float* q = queries + (size_t)qi * h->dim;
    float s = 0.f;
    for (int i = 0; i < h->dim; ++i) s += q[i] * q[i];
    const float inv = 1.0f / (sqrtf(s) + 1e-30f);
    for (int i = 0; i < h->dim; ++i) qn[i] = q[i] * inv;
    int cur = h->enter;
    float curd = dist_ip(qn, h->data + (size_t)cur * h->dim, h->dim);
    for (int l = h->maxlevel; l > 0; --l) {
      int changed = 1;
      while (changed) {
        changed = 0;
        const int* ll = links_of(h, cur, l);
        for (int j = 1; j <= ll[0]; ++j) {
          const float d = dist_ip(qn, h->data + (size_t)ll[j] * h->dim, h->dim);
          if (d < curd) { curd = d; cur = ll[j]; changed = 1; }
        }
      }
    }
    heap_t top;
    heap_init(&top, efs + 2);
    search_layer(h, qn, cur, efs, 0, &top);
    while (top.n > k) heap_pop(&top, 1);
    if (top.n < k) rc = -1;
    for (int i = top.n - 1; i >= 0; --i) {
      cand_t c = heap_pop(&top, 1);
      labels[(size_t)qi * k + i] = c.id;
      dists[(size_t)qi * k + i] = c.d;
    }
    heap_free(&top);
  }
  free(qn);
  return rc;
}
