/*
 * rrtqx_oracle.c -- CPU restatement of the RRTQX_3D geometric hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see rrtqx_oracle.h).  PARITY UNPINNED against a
 * running reference (no Julia in this image); pinned by the reference's own
 * fast-vs-naive differential pattern, its building2.txt fixture and
 * hand-evaluated known-answer vectors (tests/golden/).
 *
 * Written from the behaviour of the .jl files under /root/reference/code_RRTQx_3D; every
 * function cites the file:line it follows.  Build: oracle/Makefile
 * (gcc -O2 -ffp-contract=off -fno-fast-math).
 */
#include "rrtqx_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------- Julia semantics */

/* Base.min / Base.max on Float64 propagate NaN and order -0.0 < +0.0
 * (julia/base/math.jl, v1.0.5). */
static inline double jl_min(double x, double y) {
  if ((y < x) || (signbit(y) > signbit(x))) return isnan(x) ? x : y;
  return isnan(y) ? y : x;
}
static inline double jl_max(double x, double y) {
  if ((y > x) || (signbit(y) < signbit(x))) return isnan(x) ? x : y;
  return isnan(y) ? y : x;
}

/* ------------------------------------------------------------------ metrics */

/* DRRT_distance_functions.jl:37  sqrt(sum((x-y).^2)); sum over < 16 elements
 * is a left-to-right fold, .^2 with a literal exponent is x*x. */
double orc_euclid(const double *x, const double *y, int d) {
  double dx = x[0] - y[0];
  double s = dx * dx;
  for (int i = 1; i < d; ++i) {
    double di = x[i] - y[i];
    s = s + di * di;
  }
  return sqrt(s);
}

/* radicand only (same order); used by the sqrt-free checks in tests */
static inline double euclid_sq(const double *x, const double *y, int d) {
  double dx = x[0] - y[0];
  double s = dx * dx;
  for (int i = 1; i < d; ++i) {
    double di = x[i] - y[i];
    s = s + di * di;
  }
  return s;
}

/* DRRT_distance_functions.jl:41 */
double orc_r3sdist(const double *x, const double *y) {
  double a = euclid_sq(x, y, 3);
  double w = jl_min(fabs(x[3] - y[3]),
                    (jl_min(x[3], y[3]) + 2.0 * 3.141592653589793) -
                        jl_max(x[3], y[3]));
  return sqrt(a + w * w);
}

double orc_sqrt_threshold(double r) {
  if (isnan(r)) return r;
  if (r <= 0.0) return 0.0; /* fl(sqrt(t)) >= r for every t >= 0 */
  if (isinf(r)) return r;
  double t = r * r;
  /* walk to the smallest t with sqrt(t) >= r */
  while (sqrt(t) >= r) t = nextafter(t, -INFINITY);
  while (sqrt(t) < r) t = nextafter(t, INFINITY);
  return t;
}

/* rrtqx.jl:382 */
double orc_shrinking_ball(double delta, double ball_constant, int64_t n, int d) {
  double nn = (double)n;
  double v = ball_constant * pow(log(1.0 + nn) / nn, 1.0 / (double)d);
  return jl_min(delta, v);
}

/* ------------------------------------------------------------------ kd tree */

struct orc_kdtree {
  int d;
  int num_wraps;
  int32_t wraps[ORC_MAX_WRAPS];
  double wrap_points[ORC_MAX_WRAPS];
  int64_t n, cap;
  double *pos;       /* n x d */
  int32_t *parent;   /* kdParent   (-1: !kdParentExist) */
  int32_t *child_l;  /* kdChildL   (-1: !kdChildLExist) */
  int32_t *child_r;  /* kdChildR   (-1: !kdChildRExist) */
  int32_t *split;    /* kdSplit, 0-based */
  uint64_t dist_evals;
};

orc_kdtree *orc_kd_new(int d, int num_wraps, const int32_t *wraps,
                       const double *wrap_points) {
  if (d < 1 || d > ORC_MAX_D || num_wraps < 0 || num_wraps > ORC_MAX_WRAPS)
    return NULL;
  orc_kdtree *t = (orc_kdtree *)calloc(1, sizeof(*t));
  t->d = d;
  t->num_wraps = num_wraps;
  for (int i = 0; i < num_wraps; ++i) {
    t->wraps[i] = wraps[i];
    t->wrap_points[i] = wrap_points[i];
  }
  return t;
}

void orc_kd_free(orc_kdtree *t) {
  if (!t) return;
  free(t->pos);
  free(t->parent);
  free(t->child_l);
  free(t->child_r);
  free(t->split);
  free(t);
}

static void kd_reserve(orc_kdtree *t, int64_t need) {
  if (need <= t->cap) return;
  int64_t c = t->cap ? t->cap : 1024;
  while (c < need) c *= 2;
  t->pos = (double *)realloc(t->pos, sizeof(double) * c * t->d);
  t->parent = (int32_t *)realloc(t->parent, sizeof(int32_t) * c);
  t->child_l = (int32_t *)realloc(t->child_l, sizeof(int32_t) * c);
  t->child_r = (int32_t *)realloc(t->child_r, sizeof(int32_t) * c);
  t->split = (int32_t *)realloc(t->split, sizeof(int32_t) * c);
  t->cap = c;
}

/* kdTree_general.jl:121-170 */
int64_t orc_kd_insert(orc_kdtree *t, const double *p) {
  kd_reserve(t, t->n + 1);
  int64_t me = t->n;
  memcpy(t->pos + me * t->d, p, sizeof(double) * t->d);
  t->parent[me] = t->child_l[me] = t->child_r[me] = -1;
  if (t->n == 0) { /* :127-132 root, split = first dimension */
    t->split[me] = 0;
    t->n = 1;
    return me;
  }
  int32_t par = 0;
  for (;;) { /* :136-160 ties go right */
    int s = t->split[par];
    if (p[s] < t->pos[(int64_t)par * t->d + s]) {
      if (t->child_l[par] < 0) { t->child_l[par] = (int32_t)me; break; }
      par = t->child_l[par];
    } else {
      if (t->child_r[par] < 0) { t->child_r[par] = (int32_t)me; break; }
      par = t->child_r[par];
    }
  }
  t->parent[me] = par;
  t->split[me] = (t->split[par] == t->d - 1) ? 0 : t->split[par] + 1; /* :164-168 */
  t->n += 1;
  return me;
}

void orc_kd_insert_batch(orc_kdtree *t, const double *pos, int64_t n) {
  kd_reserve(t, t->n + n);
  for (int64_t i = 0; i < n; ++i) orc_kd_insert(t, pos + i * t->d);
}

int64_t orc_kd_size(const orc_kdtree *t) { return t->n; }
int orc_kd_dim(const orc_kdtree *t) { return t->d; }
const double *orc_kd_positions(const orc_kdtree *t) { return t->pos; }
uint64_t orc_kd_dist_evals(const orc_kdtree *t) { return t->dist_evals; }

void orc_kd_fields(const orc_kdtree *t, int32_t *parent, int32_t *child_l,
                   int32_t *child_r, int32_t *split) {
  memcpy(parent, t->parent, sizeof(int32_t) * t->n);
  memcpy(child_l, t->child_l, sizeof(int32_t) * t->n);
  memcpy(child_r, t->child_r, sizeof(int32_t) * t->n);
  memcpy(split, t->split, sizeof(int32_t) * t->n);
}

#define POS(t, i) ((t)->pos + (int64_t)(i) * (t)->d)

/* ghostPoint.jl:32-111.  One iterator per query. */
typedef struct {
  const orc_kdtree *t;
  const double *q;
  int flags[ORC_MAX_WRAPS + 1]; /* wrapDimFlags, 1-based like the reference */
  int depth;                    /* ghostTreeDepth */
  double ghost[ORC_MAX_D];      /* currentGhost */
  double closest[ORC_MAX_D];    /* closestUnwrappedPoint */
} ghost_it;

static void ghost_init(ghost_it *g, const orc_kdtree *t, const double *q) {
  g->t = t;
  g->q = q;
  memset(g->flags, 0, sizeof(g->flags));
  g->depth = t->num_wraps;
  memcpy(g->ghost, q, sizeof(double) * t->d);
  memcpy(g->closest, q, sizeof(double) * t->d);
}

/* ghostPoint.jl:60-111; returns 1 and fills out[] or 0 when exhausted */
static int ghost_next(ghost_it *g, double best_dist, double *out,
                      uint64_t *evals) {
  const orc_kdtree *t = g->t;
  for (;;) {
    while (g->depth > 0 && g->flags[g->depth] != 0) g->depth -= 1; /* :67-69 */
    if (g->depth == 0) return 0;                                   /* :71-74 */
    g->flags[g->depth] = 1;                                        /* :77 */
    int w = t->wraps[g->depth - 1];
    double wp = t->wrap_points[g->depth - 1];
    double dim_val = g->q[w];
    double dim_closest = 0.0;
    if (g->q[w] < wp / 2.0) { /* :82-89 */
      dim_val += wp;
      dim_closest += wp;
    } else {
      dim_val -= wp;
    }
    g->ghost[w] = dim_val;
    g->closest[w] = dim_closest;
    while (g->depth < t->num_wraps) { /* :96-101 */
      g->depth += 1;
      g->flags[g->depth] = 0;
      int w2 = t->wraps[g->depth - 1];
      g->ghost[w2] = g->q[w2];
      g->closest[w2] = g->ghost[w2];
    }
    if (evals) *evals += 1;
    if (orc_euclid(g->closest, g->ghost, t->d) > best_dist) continue; /* :104 */
    memcpy(out, g->ghost, sizeof(double) * t->d);
    return 1;
  }
}

/* kdTree_general.jl:254-354 */
static void kd_nearest_in_subtree(const orc_kdtree *t, int32_t root,
                                  const double *q, int32_t *best_node,
                                  double *best_dist, uint64_t *evals) {
  int32_t par = root;
  for (;;) { /* :263-280 walk down as if inserting */
    int s = t->split[par];
    if (q[s] < POS(t, par)[s]) {
      if (t->child_l[par] < 0) break;
      par = t->child_l[par];
    } else {
      if (t->child_r[par] < 0) break;
      par = t->child_r[par];
    }
  }
  double nd = orc_euclid(q, POS(t, par), t->d); /* :282-286 */
  *evals += 1;
  if (nd < *best_dist) { *best_node = par; *best_dist = nd; }

  for (;;) { /* :289-353 walk back up */
    int s = t->split[par];
    double plane = q[s] - POS(t, par)[s]; /* signed, :293 */
    if (plane > *best_dist) {            /* :295-306 */
      if (par == root) return;
      par = t->parent[par];
      continue;
    }
    if (*best_node != par) { /* :312-318 */
      nd = orc_euclid(q, POS(t, par), t->d);
      *evals += 1;
      if (nd < *best_dist) { *best_node = par; *best_dist = nd; }
    }
    if (q[s] < POS(t, par)[s] && t->child_r[par] >= 0) { /* :321-333 */
      int32_t rn = *best_node; double rd = *best_dist;
      kd_nearest_in_subtree(t, t->child_r[par], q, &rn, &rd, evals);
      if (rd < *best_dist) { *best_dist = rd; *best_node = rn; }
    } else if (POS(t, par)[s] <= q[s] && t->child_l[par] >= 0) { /* :335-345 */
      int32_t ln = *best_node; double ld = *best_dist;
      kd_nearest_in_subtree(t, t->child_l[par], q, &ln, &ld, evals);
      if (ld < *best_dist) { *best_dist = ld; *best_node = ln; }
    }
    if (par == root) return; /* :347-350 */
    par = t->parent[par];
  }
}

static int64_t kd_find_nearest_impl(const orc_kdtree *t, const double *q,
                                    double *dist_out, uint64_t *evals) {
  /* kdTree_general.jl:357-385 */
  double ldist = orc_euclid(q, POS(t, 0), t->d);
  *evals += 1;
  int32_t lnode = 0;
  kd_nearest_in_subtree(t, 0, q, &lnode, &ldist, evals);
  if (t->num_wraps > 0) {
    ghost_it g;
    ghost_init(&g, t, q);
    double gp[ORC_MAX_D];
    while (ghost_next(&g, ldist, gp, evals)) {
      double gd = orc_euclid(gp, POS(t, 0), t->d);
      *evals += 1;
      int32_t gn = 0;
      kd_nearest_in_subtree(t, 0, gp, &gn, &gd, evals);
      if (gd < ldist) { ldist = gd; lnode = gn; }
    }
  }
  if (dist_out) *dist_out = ldist;
  return lnode;
}

int64_t orc_kd_find_nearest(orc_kdtree *t, const double *q, double *dist_out) {
  if (t->n == 0) return -1; /* reference dereferences an undefined root */
  return kd_find_nearest_impl(t, q, dist_out, &t->dist_evals);
}

/* kdTree_general.jl:215-247: root first, then left subtree, then right, strict < */
int64_t orc_kd_find_nearest_naive(const orc_kdtree *t, const double *q,
                                  double *dist_out) {
  if (t->n == 0) return -1;
  /* pre-order DFS (root, L, R) with an explicit stack */
  int64_t best = -1;
  double bd = 0.0;
  int32_t *stack = (int32_t *)malloc(sizeof(int32_t) * (t->n + 1));
  int64_t sp = 0;
  stack[sp++] = 0;
  /* The recursive reference returns the subtree minimum and the caller keeps
   * its own candidate on ties, i.e. the earliest node in (root, L, R)
   * pre-order wins -- identical to a strict-< scan in that order. */
  while (sp > 0) {
    int32_t n = stack[--sp];
    double dd = orc_euclid(q, POS(t, n), t->d);
    if (best < 0 || dd < bd) { best = n; bd = dd; }
    if (t->child_r[n] >= 0) stack[sp++] = t->child_r[n];
    if (t->child_l[n] >= 0) stack[sp++] = t->child_l[n];
  }
  free(stack);
  if (dist_out) *dist_out = bd;
  return best;
}

typedef struct {
  uint8_t *marks;
  int32_t *idx;
  double *key;
  int64_t cap, len;
  int overflow;
  uint64_t evals;
} range_list;

/* addToRangeList kdTree_general.jl:765-771 */
static inline void range_push(range_list *L, int32_t node, double key) {
  if (L->marks[node]) return;
  L->marks[node] = 1;
  if (L->len < L->cap) {
    L->idx[L->len] = node;
    if (L->key) L->key[L->len] = key;
  } else {
    L->overflow = 1;
  }
  L->len += 1;
}

/* kdTree_general.jl:800-884 */
static void kd_range_in_subtree(const orc_kdtree *t, int32_t root, double range,
                                const double *q, range_list *L) {
  int32_t par = root;
  for (;;) { /* :804-822 */
    int s = t->split[par];
    if (q[s] < POS(t, par)[s]) {
      if (t->child_l[par] < 0) break;
      par = t->child_l[par];
    } else {
      if (t->child_r[par] < 0) break;
      par = t->child_r[par];
    }
  }
  double nd = orc_euclid(q, POS(t, par), t->d); /* :829-832 */
  L->evals += 1;
  if (nd < range) range_push(L, par, nd);

  for (;;) { /* :835-883 */
    int s = t->split[par];
    double plane = q[s] - POS(t, par)[s]; /* :840 signed */
    if (plane > range) {                 /* :842-853 */
      if (par == root) return;
      par = t->parent[par];
      continue;
    }
    if (!L->marks[par]) { /* :859-864 */
      nd = orc_euclid(q, POS(t, par), t->d);
      L->evals += 1;
      if (nd < range) range_push(L, par, nd);
    }
    if (q[s] < POS(t, par)[s] && t->child_r[par] >= 0) /* :867-870 */
      kd_range_in_subtree(t, t->child_r[par], range, q, L);
    else if (POS(t, par)[s] <= q[s] && t->child_l[par] >= 0) /* :871-875 */
      kd_range_in_subtree(t, t->child_l[par], range, q, L);
    if (par == root) return; /* :877-880 */
    par = t->parent[par];
  }
}

/* kdTree_general.jl:889-919 and :927-955 (identical bodies; the latter reuses L) */
static int64_t kd_range_impl(const orc_kdtree *t, double range, const double *q,
                             range_list *L) {
  double dr = orc_euclid(q, POS(t, 0), t->d); /* :895-898 root admitted with <= */
  L->evals += 1;
  if (dr <= range) range_push(L, 0, dr);
  kd_range_in_subtree(t, 0, range, q, L); /* :901 */
  if (t->num_wraps > 0) {                 /* :903-916 */
    ghost_it g;
    ghost_init(&g, t, q);
    double gp[ORC_MAX_D];
    while (ghost_next(&g, range, gp, &L->evals))
      kd_range_in_subtree(t, 0, range, gp, L);
  }
  return L->len;
}

int64_t orc_kd_find_within_range(const orc_kdtree *t, double range,
                                 const double *q, uint8_t *marks,
                                 int32_t *idx_out, double *key_out,
                                 int64_t cap, int64_t len) {
  if (t->n == 0) return 0;
  range_list L = {marks, idx_out, key_out, cap, len, 0, 0};
  kd_range_impl(t, range, q, &L);
  ((orc_kdtree *)t)->dist_evals += L.evals; /* statistic only */
  return L.overflow ? -L.len : L.len;
}

void orc_kd_empty_range_list(uint8_t *marks, const int32_t *idx, int64_t len) {
  for (int64_t i = 0; i < len; ++i) marks[idx[i]] = 0;
}

/* kdTree_general.jl:729-761: every node with dist <= range, pre-order */
int64_t orc_kd_find_within_range_naive(const orc_kdtree *t, double range,
                                       const double *q, int32_t *idx_out,
                                       double *key_out, int64_t cap) {
  int64_t len = 0;
  if (t->n == 0) return 0;
  int32_t *stack = (int32_t *)malloc(sizeof(int32_t) * (t->n + 1));
  int64_t sp = 0;
  stack[sp++] = 0;
  while (sp > 0) {
    int32_t n = stack[--sp];
    double dd = orc_euclid(q, POS(t, n), t->d);
    if (dd <= range) { /* :733 */
      if (len < cap) {
        idx_out[len] = n;
        if (key_out) key_out[len] = dd;
      }
      len += 1;
    }
    if (t->child_r[n] >= 0) stack[sp++] = t->child_r[n];
    if (t->child_l[n] >= 0) stack[sp++] = t->child_l[n];
  }
  free(stack);
  return len <= cap ? len : -len;
}

/* ------------------------------------------------------------ batch drivers */

typedef struct {
  const orc_kdtree *t;
  double range;
  const double *q;
  int64_t q0, q1;
  int32_t *counts;
  int64_t *offsets;
  int32_t *idx_out;
  double *key_out;
  int64_t cap;
  int64_t total;
  int overflow;
  /* nearest */
  int32_t *nn_idx;
  double *nn_dist;
} batch_job;

static void *range_count_worker(void *arg) {
  batch_job *j = (batch_job *)arg;
  const orc_kdtree *t = j->t;
  uint8_t *marks = (uint8_t *)calloc(t->n ? t->n : 1, 1);
  int64_t tcap = 1024;
  int32_t *tmp = (int32_t *)malloc(sizeof(int32_t) * tcap);
  for (int64_t i = j->q0; i < j->q1; ++i) {
    range_list L = {marks, tmp, NULL, tcap, 0, 0, 0};
    if (t->n) kd_range_impl(t, j->range, j->q + i * t->d, &L);
    if (L.overflow) { /* clear every mark, grow, redo */
      memset(marks, 0, t->n);
      while (tcap < L.len) tcap *= 2;
      tmp = (int32_t *)realloc(tmp, sizeof(int32_t) * tcap);
      range_list L2 = {marks, tmp, NULL, tcap, 0, 0, 0};
      kd_range_impl(t, j->range, j->q + i * t->d, &L2);
      L = L2;
    }
    j->counts[i] = (int32_t)L.len;
    j->total += L.len;
    orc_kd_empty_range_list(marks, tmp, L.len);
  }
  free(tmp);
  free(marks);
  return NULL;
}

static void *range_fill_worker(void *arg) {
  batch_job *j = (batch_job *)arg;
  const orc_kdtree *t = j->t;
  uint8_t *marks = (uint8_t *)calloc(t->n ? t->n : 1, 1);
  for (int64_t i = j->q0; i < j->q1; ++i) {
    int64_t off = j->offsets[i];
    int64_t cnt = j->counts[i];
    range_list L = {marks, j->idx_out + off, j->key_out ? j->key_out + off : NULL,
                    cnt, 0, 0, 0};
    if (t->n) kd_range_impl(t, j->range, j->q + i * t->d, &L);
    orc_kd_empty_range_list(marks, j->idx_out + off, cnt);
  }
  free(marks);
  return NULL;
}

static void run_split(void *(*fn)(void *), batch_job *proto, int nthreads) {
  int64_t n = proto->q1 - proto->q0;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if (n < nthreads) nthreads = n > 0 ? (int)n : 1;
  batch_job jobs[256];
  pthread_t th[256];
  for (int k = 0; k < nthreads; ++k) {
    jobs[k] = *proto;
    jobs[k].q0 = proto->q0 + n * k / nthreads;
    jobs[k].q1 = proto->q0 + n * (k + 1) / nthreads;
    jobs[k].total = 0;
  }
  if (nthreads == 1) {
    fn(&jobs[0]);
  } else {
    for (int k = 0; k < nthreads; ++k) pthread_create(&th[k], NULL, fn, &jobs[k]);
    for (int k = 0; k < nthreads; ++k) pthread_join(th[k], NULL);
  }
  proto->total = 0;
  for (int k = 0; k < nthreads; ++k) proto->total += jobs[k].total;
}

int64_t orc_kd_range_batch(const orc_kdtree *t, double range, const double *q,
                           int64_t q0, int64_t q1, int32_t *counts,
                           int64_t *offsets, int32_t *idx_out, double *key_out,
                           int64_t cap, int nthreads) {
  /* counts/offsets are indexed relative to q0 */
  batch_job j;
  memset(&j, 0, sizeof(j));
  j.t = t;
  j.range = range;
  j.q = q + q0 * t->d;
  j.q0 = 0;
  j.q1 = q1 - q0;
  j.counts = counts;
  run_split(range_count_worker, &j, nthreads);
  int64_t total = j.total;
  if (!idx_out) return total;
  int64_t acc = 0;
  for (int64_t i = 0; i < q1 - q0; ++i) { offsets[i] = acc; acc += counts[i]; }
  offsets[q1 - q0] = acc;
  if (acc > cap) return -1;
  j.offsets = offsets;
  j.idx_out = idx_out;
  j.key_out = key_out;
  run_split(range_fill_worker, &j, nthreads);
  return total;
}

/* Single-traversal batch driver (the timed CPU arm): every query walks the tree ONCE, as kdFindWithinRange does
 * (kdTree_general.jl:889-919), appending (node, key) to a per-thread growing buffer -- the JList pushes of the
 * reference.  The per-thread buffers are handed back as an opaque object and copied out afterwards. */
struct orc_range_lists {
  int nthreads;
  int64_t total;
  int32_t *idx[256];
  double *key[256];
  int64_t len[256];
};

typedef struct {
  batch_job j;
  int32_t *idx;
  double *key;
  int64_t cap, len;
} once_job;

static void *range_once_worker(void *arg) {
  once_job *o = (once_job *)arg;
  batch_job *j = &o->j;
  const orc_kdtree *t = j->t;
  uint8_t *marks = (uint8_t *)calloc(t->n ? t->n : 1, 1);
  o->cap = 1 << 16;
  o->len = 0;
  o->idx = (int32_t *)malloc(sizeof(int32_t) * o->cap);
  o->key = (double *)malloc(sizeof(double) * o->cap);
  for (int64_t i = j->q0; i < j->q1; ++i) {
    /* a result list cannot exceed the tree size: make room for the worst case of THIS query only when the
     * remaining space is smaller than the largest list seen so far times two (amortised doubling) */
    range_list L = {marks, o->idx + o->len, o->key + o->len, o->cap - o->len, 0, 0, 0};
    if (t->n) kd_range_impl(t, j->range, j->q + i * t->d, &L);
    if (L.overflow) { /* rare: grow and redo this one query (marks cleared first) */
      memset(marks, 0, t->n);
      while (o->cap - o->len < L.len) o->cap *= 2;
      o->idx = (int32_t *)realloc(o->idx, sizeof(int32_t) * o->cap);
      o->key = (double *)realloc(o->key, sizeof(double) * o->cap);
      range_list L2 = {marks, o->idx + o->len, o->key + o->len, o->cap - o->len, 0, 0, 0};
      kd_range_impl(t, j->range, j->q + i * t->d, &L2);
      L = L2;
    }
    j->counts[i] = (int32_t)L.len;
    orc_kd_empty_range_list(marks, o->idx + o->len, L.len);
    o->len += L.len;
    if (o->cap - o->len < 4096) { /* keep head-room so that the redo above stays rare */
      o->cap *= 2;
      o->idx = (int32_t *)realloc(o->idx, sizeof(int32_t) * o->cap);
      o->key = (double *)realloc(o->key, sizeof(double) * o->cap);
    }
  }
  free(marks);
  return NULL;
}

int64_t orc_kd_range_batch_once(const orc_kdtree *t, double range, const double *q, int64_t q0, int64_t q1,
                                int32_t *counts, int64_t *offsets, int nthreads, orc_range_lists **out) {
  int64_t n = q1 - q0;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if (n < nthreads) nthreads = n > 0 ? (int)n : 1;
  once_job jobs[256];
  pthread_t th[256];
  for (int k = 0; k < nthreads; ++k) {
    memset(&jobs[k], 0, sizeof(once_job));
    jobs[k].j.t = t;
    jobs[k].j.range = range;
    jobs[k].j.q = q + q0 * t->d;
    jobs[k].j.counts = counts;
    jobs[k].j.q0 = n * k / nthreads;
    jobs[k].j.q1 = n * (k + 1) / nthreads;
  }
  if (nthreads == 1) {
    range_once_worker(&jobs[0]);
  } else {
    for (int k = 0; k < nthreads; ++k) pthread_create(&th[k], NULL, range_once_worker, &jobs[k]);
    for (int k = 0; k < nthreads; ++k) pthread_join(th[k], NULL);
  }
  orc_range_lists *R = (orc_range_lists *)calloc(1, sizeof(orc_range_lists));
  R->nthreads = nthreads;
  for (int k = 0; k < nthreads; ++k) {
    R->idx[k] = jobs[k].idx; R->key[k] = jobs[k].key; R->len[k] = jobs[k].len;
    R->total += jobs[k].len;
  }
  if (offsets) { /* threads own consecutive query ranges, so the concatenation is in query order */
    int64_t acc = 0;
    for (int64_t i = 0; i < n; ++i) { offsets[i] = acc; acc += counts[i]; }
    offsets[n] = acc;
  }
  *out = R;
  return R->total;
}

void orc_range_lists_copy(const orc_range_lists *R, int32_t *idx, double *key) {
  int64_t acc = 0;
  for (int k = 0; k < R->nthreads; ++k) {
    if (idx) memcpy(idx + acc, R->idx[k], sizeof(int32_t) * R->len[k]);
    if (key) memcpy(key + acc, R->key[k], sizeof(double) * R->len[k]);
    acc += R->len[k];
  }
}

void orc_range_lists_free(orc_range_lists *R) {
  if (!R) return;
  for (int k = 0; k < R->nthreads; ++k) { free(R->idx[k]); free(R->key[k]); }
  free(R);
}

static void *nearest_worker(void *arg) {
  batch_job *j = (batch_job *)arg;
  uint64_t evals = 0;
  for (int64_t i = j->q0; i < j->q1; ++i) {
    double dd = 0.0;
    int64_t n = kd_find_nearest_impl(j->t, j->q + i * j->t->d, &dd, &evals);
    j->nn_idx[i] = (int32_t)n;
    if (j->nn_dist) j->nn_dist[i] = dd;
  }
  return NULL;
}

void orc_kd_nearest_batch(orc_kdtree *t, const double *q, int64_t q0,
                          int64_t q1, int32_t *idx_out, double *dist_out,
                          int nthreads) {
  if (t->n == 0) return;
  batch_job j;
  memset(&j, 0, sizeof(j));
  j.t = t;
  j.q = q + q0 * t->d;
  j.q0 = 0;
  j.q1 = q1 - q0;
  j.nn_idx = idx_out;
  j.nn_dist = dist_out;
  run_split(nearest_worker, &j, nthreads);
}

/* ------------------------------------------------- sphere world (DRRT_Q.jl) */

/* DRRT_Q.jl:1205-1210.  dot() is LinearAlgebra.dot -> OpenBLAS ddot whose
 * scalar tail is  dot = 0.0; dot += y[i]*x[i]  (x = point-start, y = end-start). */
double orc_dist_point_to_segment(const double *p, const double *s,
                                 const double *e, int d, int fma_dot) {
  double edge_len = orc_euclid(s, e, d); /* dist(startPoint,endPoint) */
  double a[ORC_MAX_D], b[ORC_MAX_D];
  for (int i = 0; i < d; ++i) { a[i] = p[i] - s[i]; b[i] = e[i] - s[i]; }
  double dot = 0.0;
  if (fma_dot) {
    for (int i = 0; i < d; ++i) dot = fma(b[i], a[i], dot);
  } else {
    for (int i = 0; i < d; ++i) dot = dot + b[i] * a[i];
  }
  double t = jl_max(0.0, jl_min(1.0, dot / edge_len));
  double c[ORC_MAX_D];
  for (int i = 0; i < d; ++i) c[i] = s[i] + t * b[i];
  return orc_euclid(p, c, d);
}

/* DRRT_Q.jl:1775-1795 (`radius == NaN` is always false, :1777) */
int orc_edge_check_sphere(const orc_sphere *ob, const double *s,
                          const double *e, double robot_radius, int fma_dot) {
  if (ob->unused || ob->life_span <= 0) return 0;
  double dist_s = orc_dist_point_to_segment(ob->pos, s, e, 3, fma_dot);
  if (dist_s > (robot_radius + ob->radius)) return 0;
  return 1;
}

/* DRRT_Q.jl:1802-1826 */
int orc_edge_check_all(const orc_sphere *obs, int64_t n_obs, int in_warmup,
                       const double *s, const double *e, double robot_radius,
                       int fma_dot) {
  if (in_warmup) return 0;
  for (int64_t i = 0; i < n_obs; ++i)
    if (orc_edge_check_sphere(&obs[i], s, e, robot_radius, fma_dot)) return 1;
  return 0;
}

/* DRRT_Q.jl:1402-1415 quickCheck2D with Wdist = euclid on [1:3]
 * (DRRT_SimpleEdge_functions.jl:61) */
static int quick_check_one(const orc_sphere *ob, const double *p) {
  if (ob->unused || ob->life_span <= 0) return 0;
  if (orc_euclid(ob->pos, p, 3) > ob->radius) return 0;
  return 1;
}

/* DRRT_Q.jl:1434-1452 */
int orc_quick_check(const orc_sphere *obs, int64_t n_obs, const double *p) {
  for (int64_t i = 0; i < n_obs; ++i)
    if (quick_check_one(&obs[i], p)) return 1;
  return 0;
}

/* DRRT_Q.jl:1463-1487 explicitPointCheck2D (== explicitPointCheck3D1 :1489-1513) */
static int point_check_one(const orc_sphere *ob, const double *p,
                           double min_dist, double robot_radius,
                           double *new_min) {
  *new_min = min_dist;
  if (ob->unused || ob->life_span <= 0) return 0;
  double this_dist = orc_euclid(ob->pos, p, 3) - robot_radius;
  if (this_dist - ob->radius > min_dist) return 0;
  this_dist = this_dist - ob->radius;
  if (this_dist < 0.0) { *new_min = 0.0; return 1; }
  *new_min = jl_min(min_dist, this_dist);
  return 0;
}

static int point_check_loop(const orc_sphere *obs, int64_t n_obs,
                            const double *p, double robot_radius,
                            double *cert_out) {
  double ret_cert = INFINITY;
  for (int64_t i = 0; i < n_obs; ++i) {
    double c;
    if (point_check_one(&obs[i], p, ret_cert, robot_radius, &c)) {
      *cert_out = 0.0;
      return 1;
    }
    if (c < ret_cert) ret_cert = c;
  }
  *cert_out = ret_cert;
  return 0;
}

/* DRRT_Q.jl:1520-1556 */
int orc_point_check(const orc_sphere *obs, int64_t n_obs, int in_warmup,
                    const double *p, double robot_radius, double *cert_out) {
  if (in_warmup) { *cert_out = INFINITY; return 0; }
  if (orc_quick_check(obs, n_obs, p)) { *cert_out = 0.0; return 1; }
  return point_check_loop(obs, n_obs, p, robot_radius, cert_out);
}

/* DRRT_Q.jl:1558-1590 */
int orc_point_check_3d(const orc_sphere *obs, int64_t n_obs, int in_warmup,
                       const double *p, double robot_radius, double *cert_out) {
  if (in_warmup) { *cert_out = INFINITY; return 0; }
  return point_check_loop(obs, n_obs, p, robot_radius, cert_out);
}

typedef struct {
  const orc_sphere *obs;
  int64_t n_obs;
  const double *pos;
  int d;
  const int32_t *src, *dst;
  int64_t e0, e1;
  double rho;
  int fma_dot;
  uint8_t *flags;
} edge_job;

static void *edge_worker(void *arg) {
  edge_job *j = (edge_job *)arg;
  for (int64_t e = j->e0; e < j->e1; ++e) {
    const double *s = j->pos + (int64_t)j->src[e] * j->d;
    const double *t = j->pos + (int64_t)j->dst[e] * j->d;
    j->flags[e] =
        (uint8_t)orc_edge_check_all(j->obs, j->n_obs, 0, s, t, j->rho, j->fma_dot);
  }
  return NULL;
}

void orc_edge_check_batch(const orc_sphere *obs, int64_t n_obs,
                          const double *pos, int d, const int32_t *src,
                          const int32_t *dst, int64_t e0, int64_t e1,
                          double robot_radius, int fma_dot, uint8_t *flags_out,
                          int nthreads) {
  int64_t n = e1 - e0;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if (n < nthreads) nthreads = n > 0 ? (int)n : 1;
  edge_job jobs[256];
  pthread_t th[256];
  for (int k = 0; k < nthreads; ++k) {
    edge_job j = {obs, n_obs, pos, d, src, dst, e0 + n * k / nthreads,
                  e0 + n * (k + 1) / nthreads, robot_radius, fma_dot, flags_out};
    jobs[k] = j;
  }
  if (nthreads == 1) {
    edge_worker(&jobs[0]);
  } else {
    for (int k = 0; k < nthreads; ++k)
      pthread_create(&th[k], NULL, edge_worker, &jobs[k]);
    for (int k = 0; k < nthreads; ++k) pthread_join(th[k], NULL);
  }
}

/* findPointsInConflictWithObstacle DRRT_Q.jl:3195-3215 (Euclidean, no theta):
 * searchRange = S.robotRadius + S.delta + ob.radius, left to right */
static int64_t conflict_candidates(const orc_kdtree *t, const orc_sphere *ob,
                                   double robot_radius, double delta,
                                   uint8_t *marks, int32_t **idx_out) {
  double search_range = (robot_radius + delta) + ob->radius;
  int64_t cap = t->n ? t->n : 1;
  int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * cap);
  double q[ORC_MAX_D] = {ob->pos[0], ob->pos[1], ob->pos[2], 0.0};
  int64_t len = orc_kd_find_within_range(t, search_range, q, marks, idx, NULL, cap, 0);
  *idx_out = idx;
  return len;
}

/* DRRT_Q.jl:3220-3290 */
int orc_obstacle_add_sweep(const orc_kdtree *t, const orc_sphere *ob_in,
                           double robot_radius, double delta,
                           const int64_t *row_ptr, const int32_t *col,
                           const int32_t *parent, int fma_dot,
                           int32_t *blocked_edges, int64_t *n_blocked,
                           int64_t cap_blocked, int32_t *orphans,
                           int64_t *n_orphans, int64_t cap_orphans,
                           int64_t *n_candidates, int64_t *n_edge_tests) {
  orc_sphere ob = *ob_in;
  ob.unused = 0; /* :3222 ob.obstacleUnused = false */
  uint8_t *marks = (uint8_t *)calloc(t->n ? t->n : 1, 1);
  int32_t *cand = NULL;
  int64_t nc = conflict_candidates(t, &ob, robot_radius, delta, marks, &cand);
  int64_t nb = 0, no = 0, ntests = 0;
  int rc = 0;
  /* popFromRangeList pops the most recently pushed first (:3233-3234) */
  for (int64_t k = nc - 1; k >= 0; --k) {
    int32_t node = cand[k];
    const double *s = POS(t, node);
    for (int64_t e = row_ptr[node]; e < row_ptr[node + 1]; ++e) { /* :3240-3252 */
      ntests += 1;
      if (orc_edge_check_sphere(&ob, s, POS(t, col[e]), robot_radius, fma_dot)) {
        if (nb < cap_blocked) blocked_edges[nb] = (int32_t)e; else rc = -1;
        nb += 1;
      }
    }
    if (parent && parent[node] >= 0) { /* :3257-3270 */
      ntests += 1;
      if (orc_edge_check_sphere(&ob, s, POS(t, parent[node]), robot_radius, fma_dot)) {
        if (no < cap_orphans) orphans[no] = node; else rc = -1;
        no += 1;
      }
    }
  }
  free(cand);
  free(marks);
  *n_blocked = nb;
  *n_orphans = no;
  if (n_candidates) *n_candidates = nc;
  if (n_edge_tests) *n_edge_tests = ntests;
  return rc;
}

/* DRRT_Q.jl:3295-3362 / DRRT.jl:3202-3268 */
int orc_obstacle_remove_sweep(const orc_kdtree *t, const orc_sphere *ob_in,
                              int check_removed_as_active,
                              const orc_sphere *others, int64_t n_others,
                              double robot_radius, double delta,
                              const int64_t *row_ptr, const int32_t *col,
                              const uint8_t *edge_dist_inf, int fma_dot,
                              int32_t *restored_edges, int64_t *n_restored,
                              int64_t cap_restored, int32_t *requeue_nodes,
                              int64_t *n_requeue, int64_t cap_requeue) {
  orc_sphere ob = *ob_in;
  uint8_t *marks = (uint8_t *)calloc(t->n ? t->n : 1, 1);
  int32_t *cand = NULL;
  int64_t nc = conflict_candidates(t, &ob, robot_radius, delta, marks, &cand);
  ob.unused = check_removed_as_active ? 0 : 1; /* DRRT_Q.jl:3301-3302 vs DRRT.jl:3267 */
  int64_t nr = 0, nq = 0;
  int rc = 0;
  for (int64_t k = nc - 1; k >= 0; --k) {
    int32_t node = cand[k];
    const double *s = POS(t, node);
    int neighbors_were_blocked = 0;
    for (int64_t e = row_ptr[node]; e < row_ptr[node + 1]; ++e) {
      const double *en = POS(t, col[e]);
      if (edge_dist_inf[e] &&
          orc_edge_check_sphere(&ob, s, en, robot_radius, fma_dot)) { /* :3319 */
        int conflicts = 0;
        for (int64_t o = 0; o < n_others; ++o) { /* :3326-3337 */
          if (orc_edge_check_sphere(&others[o], s, en, robot_radius, fma_dot)) {
            conflicts = 1;
            break;
          }
        }
        if (!conflicts) { /* :3340-3346 */
          if (nr < cap_restored) restored_edges[nr] = (int32_t)e; else rc = -1;
          nr += 1;
          neighbors_were_blocked = 1;
        }
      }
    }
    if (neighbors_were_blocked) { /* :3352-3357 */
      if (nq < cap_requeue) requeue_nodes[nq] = node; else rc = -1;
      nq += 1;
    }
  }
  free(cand);
  free(marks);
  *n_restored = nr;
  *n_requeue = nq;
  return rc;
}

/* --------------------------------------- 2-D polygon world (DRRT.jl, Otte) */

/* DRRT.jl:1060-1083 */
double orc_dist2_point_segment_2d(const double *p, const double *s,
                                  const double *e) {
  double vx = p[0] - s[0];
  double vy = p[1] - s[1];
  double ux = e[0] - s[0];
  double uy = e[1] - s[1];
  double det = vx * ux + vy * uy;
  if (det <= 0) return vx * vx + vy * vy;
  double len = ux * ux + uy * uy;
  if (det >= len) {
    double ax = e[0] - p[0], ay = e[1] - p[1];
    return ax * ax + ay * ay;
  }
  double c = ux * vy - uy * vx;
  return (c * c) / len;
}

/* DRRT.jl:1144-1202 */
double orc_segment_dist2_2d(const double *PA, const double *PB,
                            const double *QA, const double *QB) {
  int possible = 1;
  if (fabs(PB[0] - PA[0]) < .000001) { /* :1152-1157 */
    if ((QA[0] >= PA[0] && QB[0] >= PA[0]) || (QA[0] <= PA[0] && QB[0] <= PA[0]))
      possible = 0;
  } else { /* :1158-1169 */
    double m = (PB[1] - PA[1]) / (PB[0] - PA[0]);
    double diffA = (m * (QA[0] - PA[0]) + PA[1]) - QA[1];
    double diffB = (m * (QB[0] - PA[0]) + PA[1]) - QB[1];
    if ((diffA > 0.0 && diffB > 0.0) || (diffA < 0.0 && diffB < 0.0)) possible = 0;
  }
  if (possible) { /* :1172-1190 */
    if (fabs(QB[0] - QA[0]) < .000001) {
      if ((PA[0] >= QA[0] && PB[0] >= QA[0]) || (PA[0] <= QA[0] && PB[0] <= QA[0]))
        possible = 0;
    } else {
      double m = (QB[1] - QA[1]) / (QB[0] - QA[0]);
      double diffA = (m * (PA[0] - QA[0]) + QA[1]) - PA[1];
      double diffB = (m * (PB[0] - QA[0]) + QA[1]) - PB[1];
      if ((diffA > 0.0 && diffB > 0.0) || (diffA < 0.0 && diffB < 0.0)) possible = 0;
    }
  }
  if (possible) return 0.0; /* :1192-1195 */
  /* :1199-1202  min(a,b,c,d) folds left */
  double r = jl_min(orc_dist2_point_segment_2d(PA, QA, QB),
                    orc_dist2_point_segment_2d(PB, QA, QB));
  r = jl_min(r, orc_dist2_point_segment_2d(QA, PA, PB));
  r = jl_min(r, orc_dist2_point_segment_2d(QB, PA, PB));
  return r;
}

/* DRRT_data_structures.jl:229-241 */
void orc_polygon_bound(const double *poly, int32_t n_vert, double *cx,
                       double *cy, double *radius) {
  double mxx = poly[0], mnx = poly[0], mxy = poly[1], mny = poly[1];
  for (int32_t i = 1; i < n_vert; ++i) {
    mxx = jl_max(mxx, poly[2 * i]);
    mnx = jl_min(mnx, poly[2 * i]);
    mxy = jl_max(mxy, poly[2 * i + 1]);
    mny = jl_min(mny, poly[2 * i + 1]);
  }
  *cx = (mxx + mnx) / 2.0;
  *cy = (mxy + mny) / 2.0;
  double best = -INFINITY;
  for (int32_t i = 0; i < n_vert; ++i) {
    double dx = poly[2 * i] - *cx, dy = poly[2 * i + 1] - *cy;
    double s = dx * dx + dy * dy;
    best = jl_max(best, s);
  }
  *radius = sqrt(best);
}

/* DRRT.jl:1523-1578, kinds 1 and 3 without time */
int orc_edge_check_2d(const orc_obstacle2d *ob, const double *s,
                      const double *e, double radius) {
  if (ob->unused || ob->life_span <= 0) return 0;
  double d2 = orc_dist2_point_segment_2d(ob->pos, s, e); /* :1536 */
  double rr = radius + ob->radius;
  if (d2 > rr * rr) return 0; /* :1537-1539 */
  if (ob->kind == 1) return 1;
  if (ob->kind != 3) return 0;
  int32_t P = ob->n_vert;
  if (P < 2) return 0; /* :1551-1553 */
  const double *A = ob->poly + 2 * (P - 1);
  for (int32_t i = 0; i < P; ++i) { /* :1556-1578 */
    const double *B = ob->poly + 2 * i;
    if (orc_segment_dist2_2d(s, e, A, B) < radius * radius) return 1;
    A = B;
  }
  return 0;
}

/* DRRT_DubinsEdge_functions.jl:750-774 */
int orc_edge_check_dubins(const orc_obstacle2d *ob, const double *start_pos,
                          const double *end_pos, const double *traj,
                          int32_t n_traj, double robot_radius,
                          double min_turn_radius) {
  if (!orc_edge_check_2d(ob, start_pos, end_pos,
                         robot_radius + 2 * min_turn_radius))
    return 0;
  for (int32_t i = 1; i < n_traj; ++i)
    if (orc_edge_check_2d(ob, traj + 2 * (i - 1), traj + 2 * i, robot_radius))
      return 1;
  return 0;
}

/* ------------------------------- Otte / Dubins obstacle sweeps (DRRT.jl) */

/* findPointsInConflictWithObstacle, Otte generation (DRRT.jl:3048-3122), obstacle kinds 1-5, space without time:
 *   no theta:  searchRange = (rho + delta) + ob.radius            about ob.position (1x2)            :3058-3059
 *   theta   :  searchRange = ((rho + delta) + ob.radius) + pi     about [ob.x ob.y 0.0 pi]          :3061-3064
 * through kdFindWithinRange (ghost identities of the wrapped theta included). */
static int64_t conflict_candidates_2d(const orc_kdtree *t, const orc_obstacle2d *ob, int has_theta,
                                      double robot_radius, double delta, uint8_t *marks, int32_t **idx_out) {
  double search_range = (robot_radius + delta) + ob->radius;
  double q[ORC_MAX_D] = {ob->pos[0], ob->pos[1], 0.0, 0.0};
  if (has_theta) {
    search_range = search_range + 3.141592653589793;
    q[3] = 3.141592653589793;
  }
  int64_t cap = t->n ? t->n : 1;
  int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * cap);
  int64_t len = orc_kd_find_within_range(t, search_range, q, marks, idx, NULL, cap, 0);
  *idx_out = idx;
  return len;
}

/* Dubins explicitEdgeCheck of CSR item i (start node -> end node, trajectory rows traj_ptr[i]..traj_ptr[i+1]) */
static int dubins_item_check(const orc_kdtree *t, const orc_obstacle2d *ob, int32_t start, int32_t end,
                             const int64_t *traj_ptr, const double *traj_xy, int64_t item, double robot_radius,
                             double min_turn_radius) {
  return orc_edge_check_dubins(ob, POS(t, start), POS(t, end), traj_xy + 2 * traj_ptr[item],
                               (int32_t)(traj_ptr[item + 1] - traj_ptr[item]), robot_radius, min_turn_radius);
}

/* addNewObstacle, Otte generation with DubinsEdge (DRRT.jl:3127-3197) */
int orc_obstacle_add_sweep_2d(const orc_kdtree *t, const orc_obstacle2d *ob_in, int has_theta, double robot_radius,
                              double delta, double min_turn_radius, const int64_t *row_ptr, const int32_t *col,
                              const int32_t *parent, const int64_t *traj_ptr, const double *traj_xy,
                              int32_t *blocked_edges, int64_t *n_blocked, int64_t cap_blocked, int32_t *orphans,
                              int64_t *n_orphans, int64_t cap_orphans, int64_t *n_candidates,
                              int64_t *n_edge_tests) {
  orc_obstacle2d ob = *ob_in;
  ob.unused = 0; /* :3129 ob.obstacleUnused = false */
  const int64_t n_edges = row_ptr[t->n];
  uint8_t *marks = (uint8_t *)calloc(t->n ? t->n : 1, 1);
  int32_t *cand = NULL;
  int64_t nc = conflict_candidates_2d(t, &ob, has_theta, robot_radius, delta, marks, &cand);
  int64_t nb = 0, no = 0, ntests = 0;
  int rc = 0;
  for (int64_t k = nc - 1; k >= 0; --k) { /* popFromRangeList: last pushed first (:3141-3142) */
    int32_t node = cand[k];
    for (int64_t e = row_ptr[node]; e < row_ptr[node + 1]; ++e) { /* :3148-3160 */
      ntests += 1;
      if (dubins_item_check(t, &ob, node, col[e], traj_ptr, traj_xy, e, robot_radius, min_turn_radius)) {
        if (nb < cap_blocked) blocked_edges[nb] = (int32_t)e; else rc = -1;
        nb += 1;
      }
    }
    if (parent && parent[node] >= 0) { /* :3164-3177 */
      ntests += 1;
      if (dubins_item_check(t, &ob, node, parent[node], traj_ptr, traj_xy, n_edges + node, robot_radius,
                            min_turn_radius)) {
        if (no < cap_orphans) orphans[no] = node; else rc = -1;
        no += 1;
      }
    }
  }
  free(cand);
  free(marks);
  *n_blocked = nb;
  *n_orphans = no;
  if (n_candidates) *n_candidates = nc;
  if (n_edge_tests) *n_edge_tests = ntests;
  return rc;
}

/* removeObstacle, Otte generation with DubinsEdge (DRRT.jl:3202-3268): the removed obstacle is still active while
 * it is tested (obstacleUnused is set at the very end, :3267); `others` = the obstacles that pass the caller's
 * evaluation of :3238 (obOther != ob && !obstacleUnused && startTime <= timeElapsed <= startTime + lifeSpan). */
int orc_obstacle_remove_sweep_2d(const orc_kdtree *t, const orc_obstacle2d *ob, int has_theta,
                                 const orc_obstacle2d *others, int64_t n_others, double robot_radius, double delta,
                                 double min_turn_radius, const int64_t *row_ptr, const int32_t *col,
                                 const int64_t *traj_ptr, const double *traj_xy, const uint8_t *edge_dist_inf,
                                 int32_t *restored_edges, int64_t *n_restored, int64_t cap_restored,
                                 int32_t *requeue_nodes, int64_t *n_requeue, int64_t cap_requeue) {
  uint8_t *marks = (uint8_t *)calloc(t->n ? t->n : 1, 1);
  int32_t *cand = NULL;
  int64_t nc = conflict_candidates_2d(t, ob, has_theta, robot_radius, delta, marks, &cand);
  int64_t nr = 0, nq = 0;
  int rc = 0;
  for (int64_t k = nc - 1; k >= 0; --k) {
    int32_t node = cand[k];
    int neighbors_were_blocked = 0;
    for (int64_t e = row_ptr[node]; e < row_ptr[node + 1]; ++e) {
      if (edge_dist_inf[e] &&
          dubins_item_check(t, ob, node, col[e], traj_ptr, traj_xy, e, robot_radius, min_turn_radius)) { /* :3228 */
        int conflicts = 0;
        for (int64_t o = 0; o < n_others; ++o) { /* :3234-3246 */
          if (dubins_item_check(t, &others[o], node, col[e], traj_ptr, traj_xy, e, robot_radius, min_turn_radius)) {
            conflicts = 1;
            break;
          }
        }
        if (!conflicts) { /* :3249-3255 */
          if (nr < cap_restored) restored_edges[nr] = (int32_t)e; else rc = -1;
          nr += 1;
          neighbors_were_blocked = 1;
        }
      }
    }
    if (neighbors_were_blocked) { /* :3259-3264 */
      if (nq < cap_requeue) requeue_nodes[nq] = node; else rc = -1;
      nq += 1;
    }
  }
  free(cand);
  free(marks);
  *n_restored = nr;
  *n_requeue = nq;
  return rc;
}

/* ------------------------------------------------ Dubins trajectory (solver) */

/* DRRT_distance_functions.jl:62-80 */
static double right_turn_dist(const double *a, const double *b, const double *c, double r) {
  double theta = atan2(a[1] - c[1], a[0] - c[0]) - atan2(b[1] - c[1], b[0] - c[0]);
  if (theta < 0) theta = theta + 2 * 3.141592653589793;
  return theta * r;
}
static double left_turn_dist(const double *a, const double *b, const double *c, double r) {
  double theta = atan2(b[1] - c[1], b[0] - c[0]) - atan2(a[1] - c[1], a[0] - c[0]);
  if (theta < 0) theta = theta + 2 * 3.141592653589793;
  return theta * r;
}
static double dist2d(const double *a, const double *b) {
  double dx = a[0] - b[0], dy = a[1] - b[1];
  return sqrt(dx * dx + dy * dy);
}

/* appends the arc  centre + r*[cos(phi), sin(phi)]  for phi = collect(phi_start:step:phi_end)
 * (or the single angle phi_start when they are equal), DRRT_DubinsEdge_functions.jl:520-528 etc.
 * Julia's float range is TwicePrecision; here length = floor((stop-start)/step)+1 and element i =
 * start + i*step in plain binary64 -- agrees to ~1 ulp (parity by tolerance). */
static int emit_arc(const double *c, double r, double phi_start, double phi_end, double step, double *out, int n,
                    int cap) {
  int cnt = 1;
  if (phi_end != phi_start) {
    double q = (phi_end - phi_start) / step;
    cnt = (q >= 0.0) ? (int)floor(q) + 1 : 0;
  }
  for (int i = 0; i < cnt; ++i) {
    double phi = phi_start + (double)i * step;
    if (n < cap) { out[2 * n] = c[0] + r * cos(phi); out[2 * n + 1] = c[1] + r * sin(phi); }
    n += 1;
  }
  return n;
}

/* calculateTrajectory(S, edge::DubinsEdge), space without time: DRRT_DubinsEdge_functions.jl:329-709.
 * start4 / goal4 = [x y t theta].  type: 0 rsl, 1 rsr, 2 rlr, 3 lsr, 4 lsl, 5 lrl, -1 none.
 * Returns the number of trajectory points (written up to cap). */
int orc_dubins_trajectory(const double *start4, const double *goal4, double r_min, double *dist_out,
                          int32_t *type_out, double *traj_xy, int32_t cap) {
  const double PI = 3.141592653589793;
  const double il[2] = {start4[0], start4[1]}, gl[2] = {goal4[0], goal4[1]};
  const double ith = start4[3], gth = goal4[3];
  const double irc[2] = {il[0] + r_min * cos(ith - PI / 2.0), il[1] + r_min * sin(ith - PI / 2.0)};
  const double ilc[2] = {il[0] + r_min * cos(ith + PI / 2.0), il[1] + r_min * sin(ith + PI / 2.0)};
  const double grc[2] = {gl[0] + r_min * cos(gth - PI / 2.0), gl[1] + r_min * sin(gth - PI / 2.0)};
  const double glc[2] = {gl[0] + r_min * cos(gth + PI / 2.0), gl[1] + r_min * sin(gth + PI / 2.0)};
  double best = INFINITY;
  int type = -1;
  double D, v[2], R;
  /* r-s-l :373-397 */
  double rsl1[2] = {0, 0}, rsl2[2] = {0, 0};
  D = dist2d(glc, irc);
  v[0] = (glc[0] - irc[0]) / D; v[1] = (glc[1] - irc[1]) / D;
  R = -2.0 * r_min / D;
  if (!(fabs(R) > 1.0)) {
    double sq = sqrt(1.0 - R * R);
    double a = r_min * (R * v[0] + v[1] * sq), b = r_min * (R * v[1] - v[0] * sq);
    rsl1[0] = irc[0] - a; rsl2[0] = glc[0] + a;
    rsl1[1] = irc[1] - b; rsl2[1] = glc[1] + b;
    double len = right_turn_dist(il, rsl1, irc, r_min) + dist2d(rsl2, rsl1) + left_turn_dist(rsl2, gl, glc, r_min);
    if (best > len) { best = len; type = 0; }
  }
  /* r-s-r :400-415 */
  double rsr1[2], rsr2[2];
  D = dist2d(grc, irc);
  v[0] = (grc[0] - irc[0]) / D; v[1] = (grc[1] - irc[1]) / D;
  rsr1[0] = -r_min * v[1] + irc[0]; rsr2[0] = -r_min * v[1] + grc[0];
  rsr1[1] = r_min * v[0] + irc[1];  rsr2[1] = r_min * v[0] + grc[1];
  {
    double len = right_turn_dist(il, rsr1, irc, r_min) + dist2d(rsr2, rsr1) + right_turn_dist(rsr2, gl, grc, r_min);
    if (best > len) { best = len; type = 1; }
  }
  /* r-l-r :418-433 (uses D, v of the r-s-r block) */
  double rlr_rl[2] = {NAN, NAN}, rlr_lr[2] = {NAN, NAN}, rlr_c[2] = {NAN, NAN};
  if (D < 4.0 * r_min) {
    double theta = -acos(D / (4 * r_min)) + atan2(v[1], v[0]);
    rlr_c[0] = irc[0] + 2 * r_min * cos(theta); rlr_c[1] = irc[1] + 2 * r_min * sin(theta);
    rlr_rl[0] = (rlr_c[0] + irc[0]) / 2.0; rlr_rl[1] = (rlr_c[1] + irc[1]) / 2.0;
    rlr_lr[0] = (rlr_c[0] + grc[0]) / 2.0; rlr_lr[1] = (rlr_c[1] + grc[1]) / 2.0;
    double len = right_turn_dist(il, rlr_rl, irc, r_min) + left_turn_dist(rlr_rl, rlr_lr, rlr_c, r_min) +
                 right_turn_dist(rlr_lr, gl, grc, r_min);
    if (best > len) { best = len; type = 2; }
  }
  /* l-s-r :436-460 */
  double lsr1[2] = {0, 0}, lsr2[2] = {0, 0};
  D = dist2d(grc, ilc);
  v[0] = (grc[0] - ilc[0]) / D; v[1] = (grc[1] - ilc[1]) / D;
  R = 2.0 * r_min / D;
  if (!(fabs(R) > 1)) {
    double sq = sqrt(1 - R * R);
    double a = R * v[0] + v[1] * sq, b = R * v[1] - v[0] * sq;
    lsr1[0] = ilc[0] + a * r_min; lsr2[0] = grc[0] - a * r_min;
    lsr1[1] = ilc[1] + b * r_min; lsr2[1] = grc[1] - b * r_min;
    double len = left_turn_dist(il, lsr1, ilc, r_min) + dist2d(lsr2, lsr1) + right_turn_dist(lsr2, gl, grc, r_min);
    if (best > len) { best = len; type = 3; }
  }
  /* l-s-l :463-478 */
  double lsl1[2], lsl2[2];
  D = dist2d(glc, ilc);
  v[0] = (glc[0] - ilc[0]) / D; v[1] = (glc[1] - ilc[1]) / D;
  lsl1[0] = r_min * v[1] + ilc[0];  lsl2[0] = r_min * v[1] + glc[0];
  lsl1[1] = -r_min * v[0] + ilc[1]; lsl2[1] = -r_min * v[0] + glc[1];
  {
    double len = left_turn_dist(il, lsl1, ilc, r_min) + dist2d(lsl2, lsl1) + left_turn_dist(lsl2, gl, glc, r_min);
    if (best > len) { best = len; type = 4; }
  }
  /* l-r-l :481-499 */
  double lrl_lr[2] = {NAN, NAN}, lrl_rl[2] = {NAN, NAN}, lrl_c[2] = {NAN, NAN};
  if (D < 4.0 * r_min) {
    double theta = acos(D / (4 * r_min)) + atan2(v[1], v[0]);
    lrl_c[0] = ilc[0] + 2.0 * r_min * cos(theta); lrl_c[1] = ilc[1] + 2.0 * r_min * sin(theta);
    lrl_lr[0] = (lrl_c[0] + ilc[0]) / 2.0; lrl_lr[1] = (lrl_c[1] + ilc[1]) / 2.0;
    lrl_rl[0] = (lrl_c[0] + glc[0]) / 2.0; lrl_rl[1] = (lrl_c[1] + glc[1]) / 2.0;
    double len = left_turn_dist(il, lrl_lr, ilc, r_min) + right_turn_dist(lrl_lr, lrl_rl, lrl_c, r_min) +
                 left_turn_dist(lrl_rl, gl, glc, r_min);
    if (best > len) { best = len; type = 5; }
  }
  *dist_out = best;
  *type_out = type;
  if (type < 0) return 0;
  static const char *names[6] = {"rsl", "rsr", "rlr", "lsr", "lsl", "lrl"};
  const char *nm = names[type];
  const double dphi = .1;
  int n = 0;
  /* first part :508-548 */
  if (nm[0] == 'r') {
    const double *p = type == 0 ? rsl1 : (type == 1 ? rsr1 : rlr_rl);
    double ps = atan2(il[1] - irc[1], il[0] - irc[0]), pe = atan2(p[1] - irc[1], p[0] - irc[0]);
    if (pe > ps) pe = pe - 2.0 * PI;
    n = emit_arc(irc, r_min, ps, pe, -dphi, traj_xy, n, cap);
  } else {
    const double *p = type == 4 ? lsl1 : (type == 3 ? lsr1 : lrl_lr);
    double ps = atan2(il[1] - ilc[1], il[0] - ilc[0]), pe = atan2(p[1] - ilc[1], p[0] - ilc[0]);
    if (pe < ps) pe = pe + 2.0 * PI;
    n = emit_arc(ilc, r_min, ps, pe, dphi, traj_xy, n, cap);
  }
  /* second part :552-603 */
  if (nm[1] == 's') {
    const double *p1 = type == 3 ? lsr1 : (type == 4 ? lsl1 : (type == 1 ? rsr1 : rsl1));
    const double *p2 = type == 3 ? lsr2 : (type == 4 ? lsl2 : (type == 1 ? rsr2 : rsl2));
    if (n < cap) { traj_xy[2 * n] = p1[0]; traj_xy[2 * n + 1] = p1[1]; }
    n += 1;
    if (n < cap) { traj_xy[2 * n] = p2[0]; traj_xy[2 * n + 1] = p2[1]; }
    n += 1;
  } else if (nm[1] == 'r') { /* lrl */
    double ps = atan2(lrl_lr[1] - lrl_c[1], lrl_lr[0] - lrl_c[0]), pe = atan2(lrl_rl[1] - lrl_c[1], lrl_rl[0] - lrl_c[0]);
    if (pe > ps) pe = pe - 2.0 * PI;
    n = emit_arc(lrl_c, r_min, ps, pe, -dphi, traj_xy, n, cap);
  } else { /* rlr */
    double ps = atan2(rlr_rl[1] - rlr_c[1], rlr_rl[0] - rlr_c[0]), pe = atan2(rlr_lr[1] - rlr_c[1], rlr_lr[0] - rlr_c[0]);
    if (pe < ps) pe = pe + 2.0 * PI;
    n = emit_arc(rlr_c, r_min, ps, pe, dphi, traj_xy, n, cap);
  }
  /* third part :606-655 */
  if (nm[2] == 'r') {
    const double *p = type == 1 ? rsr2 : (type == 3 ? lsr2 : rlr_lr);
    double ps = atan2(p[1] - grc[1], p[0] - grc[0]), pe = atan2(gl[1] - grc[1], gl[0] - grc[0]);
    if (pe > ps) pe = pe - 2.0 * PI;
    n = emit_arc(grc, r_min, ps, pe, -dphi, traj_xy, n, cap);
  } else {
    const double *p = type == 4 ? lsl2 : (type == 0 ? rsl2 : lrl_rl);
    double ps = atan2(p[1] - glc[1], p[0] - glc[0]), pe = atan2(gl[1] - glc[1], gl[0] - glc[0]);
    if (pe < ps) pe = pe + 2.0 * PI;
    n = emit_arc(glc, r_min, ps, pe, dphi, traj_xy, n, cap);
  }
  return n;
}

/* saturate, DubinsEdge: DRRT_DubinsEdge_functions.jl:70-95 is read below */
void orc_saturate_dubins(double *new_point, const double *closest, double delta) {
  const double PI = 3.141592653589793;
  double this_dist = orc_r3sdist(new_point, closest); /* dist = R3SDist, :41 */
  if (this_dist > delta) {
    for (int i = 0; i < 3; ++i) /* :76 */
      new_point[i] = closest[i] + (new_point[i] - closest[i]) * delta / this_dist;
    if (fabs(new_point[3] - closest[3]) < PI) { /* :79-81 */
      new_point[3] = closest[3] + (new_point[3] - closest[3]) * delta / this_dist;
    } else { /* :82-93 */
      new_point[3] = (new_point[3] < PI ? new_point[3] + 2 * PI : new_point[3] - 2 * PI);
      new_point[3] = closest[3] + (new_point[3] - closest[3]) * delta / this_dist;
      new_point[3] = jl_max(jl_min(new_point[3], 2 * PI), 0.0);
    }
  }
}
