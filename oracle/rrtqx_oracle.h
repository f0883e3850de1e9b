/*
 * rrtqx_oracle.h -- CPU restatement of the RRTQX_3D geometric hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (rrtqx_3d_b200/, the
 * C-ABI library, the Julia module) may include, link or call this.  It is used
 * by tests/, by __graft_entry__.smoke() as the checker, and by bench.py's
 * cpu_baseline / --impl reference legs as the timed CPU arm.
 *
 * PARITY UNPINNED against a *running* reference: the reference is Julia 1.0.5,
 * there is no Julia here and the reference ships no executable tests or golden
 * vectors for this path.  What pins this file instead:
 *   (i)   the reference's own fast-vs-naive differential pattern
 *         (kdTree_general.jl:1039-1148) re-run on this restatement,
 *   (ii)  the reference fixture environments/building2.txt (31 spheres),
 *   (iii) hand-evaluated known-answer vectors in tests/golden/.
 *
 * All citations are file:line under /root/reference/code_RRTQx_3D/.
 * Arithmetic: IEEE binary64, every + - * / sqrt individually rounded, written
 * order, never fused (compile with -ffp-contract=off).
 */
#ifndef RRTQX_ORACLE_H
#define RRTQX_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_D 4
#define ORC_MAX_WRAPS 4

/* ------------------------------------------------------------------ metrics */
/* DRRT_distance_functions.jl:37  euclidianDist, first d coordinates */
double orc_euclid(const double *x, const double *y, int d);
/* DRRT_distance_functions.jl:41  R3SDist on [x y t theta] */
double orc_r3sdist(const double *x, const double *y);
/* T*(r) = min{ t : fl(sqrt(t)) >= r }  (SURVEY appendix A4); helper for tests */
double orc_sqrt_threshold(double r);
/* rrtqx.jl:382 shrinking ball; libm log/pow, NOT Julia's -- host-side constant */
double orc_shrinking_ball(double delta, double ball_constant, int64_t n, int d);

/* ------------------------------------------------------------------ kd tree */
typedef struct orc_kdtree orc_kdtree;

/* KDTree{T}(d,f[,wraps,wrapPoints])  kdTree_general.jl:94-112.
 * wraps are 0-based dimension indices here (Julia's are 1-based). */
orc_kdtree *orc_kd_new(int d, int num_wraps, const int32_t *wraps,
                       const double *wrap_points);
void orc_kd_free(orc_kdtree *t);
/* kdInsert kdTree_general.jl:121-170; returns the 0-based node index */
int64_t orc_kd_insert(orc_kdtree *t, const double *pos);
void orc_kd_insert_batch(orc_kdtree *t, const double *pos, int64_t n);
int64_t orc_kd_size(const orc_kdtree *t);
int orc_kd_dim(const orc_kdtree *t);
const double *orc_kd_positions(const orc_kdtree *t); /* n x d row-major */
/* kd fields per node: parent/childL/childR = -1 when absent, split 0-based */
void orc_kd_fields(const orc_kdtree *t, int32_t *parent, int32_t *child_l,
                   int32_t *child_r, int32_t *split);
/* number of distanceFunction evaluations since creation (cost statistic);
 * only counted by the single-threaded entry points */
uint64_t orc_kd_dist_evals(const orc_kdtree *t);

/* kdFindNearest kdTree_general.jl:357-385 (incl. ghost identities) */
int64_t orc_kd_find_nearest(orc_kdtree *t, const double *q, double *dist_out);
/* kdFindNearestNaive :245 */
int64_t orc_kd_find_nearest_naive(const orc_kdtree *t, const double *q,
                                  double *dist_out);

/* kdFindWithinRange :889-919 / kdFindMoreWithinRange :927-955.
 * marks = the per-node inHeap flags (uint8[n], caller-owned so that several
 * threads can query one immutable tree).  Entries are appended in PUSH order
 * at idx_out[len..], key_out[len..]; the Julia JList pops them in reverse.
 * Pass len = 0 and cleared marks for kdFindWithinRange, the previous length
 * and un-cleared marks for kdFindMoreWithinRange.  Returns the new length, or
 * -(needed) if cap is too small (marks are then left in an undefined state). */
int64_t orc_kd_find_within_range(const orc_kdtree *t, double range,
                                 const double *q, uint8_t *marks,
                                 int32_t *idx_out, double *key_out,
                                 int64_t cap, int64_t len);
/* emptyRangeList :782-787 : clears marks of the listed nodes */
void orc_kd_empty_range_list(uint8_t *marks, const int32_t *idx, int64_t len);
/* kdFindWithinRangeNaive :754-761 (note: tests <= range for every node) */
int64_t orc_kd_find_within_range_naive(const orc_kdtree *t, double range,
                                       const double *q, int32_t *idx_out,
                                       double *key_out, int64_t cap);

/* Batch drivers used for fixtures and for the timed CPU baseline.  Queries
 * [q0,q1) of an nq x d row-major array.  counts[i] receives the list length of
 * query q0+i.  If idx_out != NULL results are appended CSR-style and
 * offsets[i] (int64, nq+1 entries relative to q0) are written; cap as above.
 * nthreads > 1 splits the query range over pthreads. Returns total results or
 * -1 on overflow. */
int64_t orc_kd_range_batch(const orc_kdtree *t, double range, const double *q,
                           int64_t q0, int64_t q1, int32_t *counts,
                           int64_t *offsets, int32_t *idx_out, double *key_out,
                           int64_t cap, int nthreads);
/* Single-traversal form of the same batch (what bench.py's CPU arm times): one walk per query, results in
 * per-thread growing buffers; counts[i] / offsets[i] (nq+1) as above.  Returns the total; the lists are copied
 * out (query order) with orc_range_lists_copy and released with orc_range_lists_free. */
typedef struct orc_range_lists orc_range_lists;
int64_t orc_kd_range_batch_once(const orc_kdtree *t, double range, const double *q, int64_t q0, int64_t q1,
                                int32_t *counts, int64_t *offsets, int nthreads, orc_range_lists **out);
void orc_range_lists_copy(const orc_range_lists *R, int32_t *idx, double *key);
void orc_range_lists_free(orc_range_lists *R);
void orc_kd_nearest_batch(orc_kdtree *t, const double *q, int64_t q0,
                          int64_t q1, int32_t *idx_out, double *dist_out,
                          int nthreads);

/* ------------------------------------------------- sphere world (DRRT_Q.jl) */
/* SphereObstacle, DRRT_data_structures.jl:267-307, flattened */
typedef struct {
  double pos[3];
  double radius;
  double start_time;
  double life_span;
  uint8_t unused; /* obstacleUnused */
  uint8_t pad[7];
} orc_sphere;

/* distancePointToSegment DRRT_Q.jl:1205-1210 (d = array length, 3 in QX).
 * fma_dot != 0 evaluates the ddot tail with fused multiply-adds. */
double orc_dist_point_to_segment(const double *p, const double *s,
                                 const double *e, int d, int fma_dot);
/* explicitEdgeCheck3D DRRT_Q.jl:1775-1795 */
int orc_edge_check_sphere(const orc_sphere *ob, const double *s,
                          const double *e, double robot_radius, int fma_dot);
/* explicitEdgeCheck(C,edge) DRRT_Q.jl:1802-1826: OR over obs[0..n) in order */
int orc_edge_check_all(const orc_sphere *obs, int64_t n_obs, int in_warmup,
                       const double *s, const double *e, double robot_radius,
                       int fma_dot);
/* quickCheck DRRT_Q.jl:1434-1452 (via quickCheck2D :1402) */
int orc_quick_check(const orc_sphere *obs, int64_t n_obs, const double *p);
/* explicitPointCheck DRRT_Q.jl:1520-1556: returns collide flag, *cert_out */
int orc_point_check(const orc_sphere *obs, int64_t n_obs, int in_warmup,
                    const double *p, double robot_radius, double *cert_out);
/* explicitPointCheck3D DRRT_Q.jl:1558-1590 (no quickCheck pass) */
int orc_point_check_3d(const orc_sphere *obs, int64_t n_obs, int in_warmup,
                       const double *p, double robot_radius, double *cert_out);

/* Batched drivers (CPU baseline).  pos: n x d row-major node positions,
 * edges given as (src,dst) index pairs.  flags_out[i] in {0,1}. */
void orc_edge_check_batch(const orc_sphere *obs, int64_t n_obs,
                          const double *pos, int d, const int32_t *src,
                          const int32_t *dst, int64_t e0, int64_t e1,
                          double robot_radius, int fma_dot, uint8_t *flags_out,
                          int nthreads);

/* addNewObstacle DRRT_Q.jl:3220-3290 restricted to its geometric decisions.
 * Graph given as out-edge CSR (row = start node, edge ids are CSR positions,
 * listing InitialNeighborListOut then rrtNeighborsOut as the reference
 * iterator does, DRRT_Q.jl:2408-2431) and parent[] (-1 = rrtParentUsed false).
 * For obstacle `ob`: candidate nodes = kdFindWithinRange(KD,(rho+delta)+R,
 * ob.pos) (:3203-3204); every out-edge of a candidate that collides is
 * reported in blocked_edges (ids, possibly repeated when listed twice), every
 * candidate whose parent edge collides in orphans.  Returns 0, counts through
 * the pointers; -1 on capacity overflow. */
int orc_obstacle_add_sweep(const orc_kdtree *t, const orc_sphere *ob,
                           double robot_radius, double delta,
                           const int64_t *row_ptr, const int32_t *col,
                           const int32_t *parent, int fma_dot,
                           int32_t *blocked_edges, int64_t *n_blocked,
                           int64_t cap_blocked, int32_t *orphans,
                           int64_t *n_orphans, int64_t cap_orphans,
                           int64_t *n_candidates, int64_t *n_edge_tests);

/* removeObstacle DRRT_Q.jl:3295-3362, geometric decisions only.  edge_dist_inf
 * = per-edge flag "edge.dist == Inf".  `others` = the obstacles that pass the
 * reference's "other obstacle" predicate (:3330), evaluated by the caller.
 * check_removed_as_active: 1 = Otte behaviour (DRRT.jl:3202-3268, the removed
 * obstacle is still active while tested), 0 = QX behaviour (it was disabled
 * first, so nothing is ever restored). */
int orc_obstacle_remove_sweep(const orc_kdtree *t, const orc_sphere *ob,
                              int check_removed_as_active,
                              const orc_sphere *others, int64_t n_others,
                              double robot_radius, double delta,
                              const int64_t *row_ptr, const int32_t *col,
                              const uint8_t *edge_dist_inf, int fma_dot,
                              int32_t *restored_edges, int64_t *n_restored,
                              int64_t cap_restored, int32_t *requeue_nodes,
                              int64_t *n_requeue, int64_t cap_requeue);

/* --------------------------------------- 2-D polygon world (DRRT.jl, Otte) */
/* distanceSqrdPointToSegment DRRT.jl:1060-1083 */
double orc_dist2_point_segment_2d(const double *p, const double *s,
                                  const double *e);
/* segmentDistSqrd DRRT.jl:1144-1202 */
double orc_segment_dist2_2d(const double *pa, const double *pb,
                            const double *qa, const double *qb);
/* Obstacle kinds 1 (ball) and 3 (polygon), DRRT_data_structures.jl:135-265 */
typedef struct {
  int32_t kind;   /* 1 or 3 */
  int32_t n_vert; /* polygon rows (kind 3) */
  double pos[2];  /* bounding-circle centre */
  double radius;  /* bounding-circle radius */
  double life_span;
  uint8_t unused;
  uint8_t pad[7];
  const double *poly; /* n_vert x 2 row-major */
} orc_obstacle2d;
/* bounding circle of a polygon, DRRT_data_structures.jl:229-241 */
void orc_polygon_bound(const double *poly, int32_t n_vert, double *cx,
                       double *cy, double *radius);
/* explicitEdgeCheck2D DRRT.jl:1523-1578 (kinds 1,3; no time dimension) */
int orc_edge_check_2d(const orc_obstacle2d *ob, const double *s,
                      const double *e, double radius);
/* Dubins explicitEdgeCheck DRRT_DubinsEdge_functions.jl:750-774.
 * traj: n_traj x 2 row-major (edge.trajectory[:,1:2]). */
int orc_edge_check_dubins(const orc_obstacle2d *ob, const double *start_pos,
                          const double *end_pos, const double *traj,
                          int32_t n_traj, double robot_radius,
                          double min_turn_radius);

/* addNewObstacle / removeObstacle of the Otte generation with DubinsEdge (DRRT.jl:3048-3268), geometric
 * decisions only.  Candidates: kdFindWithinRange(KD, ((rho+delta)+ob.radius)+pi, [ob.x ob.y 0.0 pi]) when
 * has_theta (d = 4 tree with the theta wrap), else (rho+delta)+ob.radius about ob.position.  Graph as in the
 * sphere sweeps (out-edge CSR; edge id = CSR position; parent[] = end node of rrtParentEdge or -1).  Trajectories
 * (edge.trajectory[:,1:2]) as one CSR over the ITEMS: item e < n_edges is CSR edge e, item n_edges + v is the
 * parent edge of node v (traj_ptr has n_edges + n_nodes + 1 entries).  Edge test: orc_edge_check_dubins. */
int orc_obstacle_add_sweep_2d(const orc_kdtree *t, const orc_obstacle2d *ob, int has_theta, double robot_radius,
                              double delta, double min_turn_radius, const int64_t *row_ptr, const int32_t *col,
                              const int32_t *parent, const int64_t *traj_ptr, const double *traj_xy,
                              int32_t *blocked_edges, int64_t *n_blocked, int64_t cap_blocked, int32_t *orphans,
                              int64_t *n_orphans, int64_t cap_orphans, int64_t *n_candidates,
                              int64_t *n_edge_tests);
int orc_obstacle_remove_sweep_2d(const orc_kdtree *t, const orc_obstacle2d *ob, int has_theta,
                                 const orc_obstacle2d *others, int64_t n_others, double robot_radius, double delta,
                                 double min_turn_radius, const int64_t *row_ptr, const int32_t *col,
                                 const int64_t *traj_ptr, const double *traj_xy, const uint8_t *edge_dist_inf,
                                 int32_t *restored_edges, int64_t *n_restored, int64_t cap_restored,
                                 int32_t *requeue_nodes, int64_t *n_requeue, int64_t cap_requeue);

/* Dubins calculateTrajectory, space without time (DRRT_DubinsEdge_functions.jl:329-709): six-word
 * solver + arc discretisation at 0.1 rad.  start4/goal4 = [x y t theta].  type: 0 rsl, 1 rsr, 2 rlr,
 * 3 lsr, 4 lsl, 5 lrl, -1 none.  Returns the number of trajectory points (x,y rows; written up to
 * cap).  Uses the C libm, not Julia's: parity with the reference is by tolerance (1e-9), SURVEY A14. */
int orc_dubins_trajectory(const double *start4, const double *goal4, double r_min, double *dist_out,
                          int32_t *type_out, double *traj_xy, int32_t cap);
/* saturate, DubinsEdge (DRRT_DubinsEdge_functions.jl:70-95), in place on new_point[4] */
void orc_saturate_dubins(double *new_point, const double *closest, double delta);

#ifdef __cplusplus
}
#endif
#endif
