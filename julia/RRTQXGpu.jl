# RRTQXGpu.jl -- Julia FFI module for librrtqx_b200.so (include/rrtqx_b200.h).
#
# Drop-in for the geometric inner loop of RRTQX_3D: it re-defines, on top of `ccall`s into the
# hand-written sm_100a library, the generic functions the planner calls
#
#   kdInsert, kdFindNearest, kdFindWithinRange, kdFindMoreWithinRange        (kdTree_general.jl:121,357,889,927)
#   explicitEdgeCheck(S, edge[, ob]), explicitPointCheck, explicitNodeCheck  (DRRT_Q.jl:1802,1520,1594; DRRT_SimpleEdge_functions.jl:210)
#   findPointsInConflictWithObstacle, addNewObstacle, removeObstacle          (DRRT_Q.jl:3195,3220,3295)
#
# for a tree of type `GpuKDTree{T}`; popFromRangeList / emptyRangeList (kdTree_general.jl:774-787) are
# unchanged because the results are rebuilt as ordinary JLists with inHeap marks.
#
# Usage in an experiment script (after the reference's own includes):
#     include("RRTQXGpu.jl"); using .RRTQXGpu
#     ctx = RRTQXGpu.Context(0)
#     KD  = RRTQXGpu.GpuKDTree{RRTNode{Float64}}(ctx, S.d, KDdist)          # instead of KDTree{...}(S.d, KDdist)
#
# NOTE: there is no Julia in the build image of this repository, so this file is written blind against
# Julia 1.0 syntax and kept thin; all logic lives behind the C ABI, where it is tested (tests/, via the
# Python ctypes twin rrtqx_3d_b200/_abi.py which binds the same symbols with the same argument order).
# There is no CPU fallback: a missing library or GPU is an error().

module RRTQXGpu

export Context, GpuKDTree, GpuObstacles, GpuEdges,
       kdInsert, kdInsertBatch, kdFindNearest, kdFindWithinRange, kdFindMoreWithinRange,
       kdFindWithinRangeBatch, kdFindNearestBatch,
       explicitEdgeCheck, explicitEdgeCheckBatch, explicitPointCheck, explicitPointCheck3D,
       explicitNodeCheck, explicitNodeCheck3D,
       findPointsInConflictWithObstacle, addNewObstacle, removeObstacle, syncObstacles!, syncEdges!

const LIB = get(ENV, "RRTQX_B200_LIB", joinpath(@__DIR__, "..", "rrtqx_3d_b200", "librrtqx_b200.so"))

const RANGE_WANT_DIST  = UInt32(1)
const RANGE_COUNT_ONLY = UInt32(2)
const CHECK_FMA_DOT       = UInt32(1)
const CHECK_IGNORE_ACTIVE = UInt32(2)
const CHECK_QUICK_PASS    = UInt32(4)
const SWEEP_REMOVED_INACTIVE = UInt32(16)

# ------------------------------------------------------------------ context
mutable struct Context
  h::Ptr{Cvoid}
  function Context(device::Integer = 0)
    isfile(LIB) || error("librrtqx_b200.so not found at $(LIB); build it first (no CPU fallback)")
    h = Ref{Ptr{Cvoid}}(C_NULL)
    st = ccall((:rrtqx_ctx_create, LIB), Int32, (Int32, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), device, C_NULL, h)
    st == 0 || error("rrtqx_ctx_create: " * unsafe_string(ccall((:rrtqx_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    c = new(h[])
    finalizer(x -> ccall((:rrtqx_ctx_destroy, LIB), Int32, (Ptr{Cvoid},), x.h), c)
    return c
  end
end

function check(ctx::Context, st::Int32)
  st == 0 && return nothing
  error("rrtqx status $(st): " * unsafe_string(ccall((:rrtqx_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.h)))
end

# --------------------------------------------------------------------- tree
# T is the node type (RRTNode{Float64}); nodes keep their own kd fields and are owned by Julia.
mutable struct GpuKDTree{T}
  ctx::Context
  h::Ptr{Cvoid}
  d::Int
  distanceFunction::Function      # kept for API compatibility; the device metric is KDdist = euclidianDist
  treeSize::Int
  numWraps::Int
  wraps::Array{Int}
  wrapPoints::Array{Float64}
  root::T
  nodes::Vector{T}                # device index (0-based) + 1 -> node
  index::IdDict{Any,Int32}        # node -> device index (RRTNode has no integer id)
  res::Base.RefValue{Ptr{Cvoid}}  # reusable rrtqx_range_result

  function GpuKDTree{T}(ctx::Context, d::Int, f::Function, wraps::Array{Int} = Int[],
                        wrapPoints::Array{Float64} = Float64[]) where {T}
    h = Ref{Ptr{Cvoid}}(C_NULL)
    w0 = Int32[w - 1 for w in wraps]   # the C ABI uses 0-based dimensions
    check(ctx, ccall((:rrtqx_tree_create, LIB), Int32,
                     (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
                     ctx.h, d, length(w0), w0, wrapPoints, h))
    t = new{T}(ctx, h[], d, f, 0, length(wraps), wraps, wrapPoints)
    t.nodes = Vector{T}()
    t.index = IdDict{Any,Int32}()
    t.res = Ref{Ptr{Cvoid}}(C_NULL)
    finalizer(t) do x
      x.res[] != C_NULL && ccall((:rrtqx_range_result_destroy, LIB), Int32, (Ptr{Cvoid},), x.res[])
      ccall((:rrtqx_tree_destroy, LIB), Int32, (Ptr{Cvoid},), x.h)
    end
    return t
  end
end

# keeps kdParent / kdChildL / kdChildR / kdSplit of the Julia nodes populated (saveRRTTree etc. walk them)
function refreshKdFields!(tree::GpuKDTree, first::Int, count::Int)
  parent = Vector{Int32}(undef, count); cl = similar(parent); cr = similar(parent); sp = similar(parent)
  check(tree.ctx, ccall((:rrtqx_tree_kd_fields, LIB), Int32,
                        (Ptr{Cvoid}, Int64, Int64, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
                        tree.h, first, count, parent, cl, cr, sp))
  for k = 1:count
    n = tree.nodes[first + k]
    n.kdSplit = sp[k] + 1
    if parent[k] >= 0
      p = tree.nodes[parent[k] + 1]
      n.kdParent = p
      n.kdParentExist = true
      if n.position[p.kdSplit] < p.position[p.kdSplit]
        p.kdChildL = n; p.kdChildLExist = true
      else
        p.kdChildR = n; p.kdChildRExist = true
      end
    end
  end
end

# kdInsert (kdTree_general.jl:121-170)
function kdInsert(tree::GpuKDTree{T}, node::T) where {T}
  node.kdInTree && return
  node.kdInTree = true
  idx = Ref{Int32}(0)
  pos = vec(Array{Float64}(node.position))
  GC.@preserve pos check(tree.ctx, ccall((:rrtqx_tree_insert, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ref{Int32}),
                                         tree.h, pos, idx))
  push!(tree.nodes, node)
  tree.index[node] = idx[]
  if tree.treeSize == 0
    tree.root = node
  end
  tree.treeSize += 1
  refreshKdFields!(tree, Int(idx[]), 1)
  return
end

function kdInsertBatch(tree::GpuKDTree{T}, nodes::Vector{T}) where {T}
  new = [n for n in nodes if !n.kdInTree]
  isempty(new) && return
  pos = Matrix{Float64}(undef, tree.d, length(new))       # column-major d x n == row-major n x d for C
  for (k, n) in enumerate(new); pos[:, k] = vec(n.position); end
  first = Ref{Int32}(0)
  GC.@preserve pos check(tree.ctx, ccall((:rrtqx_tree_insert_batch, LIB), Int32,
                                         (Ptr{Cvoid}, Ptr{Float64}, Int64, Ref{Int32}), tree.h, pos, length(new), first))
  for (k, n) in enumerate(new)
    n.kdInTree = true
    push!(tree.nodes, n)
    tree.index[n] = first[] + Int32(k - 1)
  end
  if tree.treeSize == 0
    tree.root = new[1]
  end
  tree.treeSize += length(new)
  refreshKdFields!(tree, Int(first[]), length(new))
  return
end

# kdFindNearest (kdTree_general.jl:357-385) -> (node, dist)
function kdFindNearest(tree::GpuKDTree, queryPoint::Array{Float64})
  tree.treeSize == 0 && error("kdFindNearest on an empty tree")
  q = vec(Array{Float64}(queryPoint)); idx = Vector{Int32}(undef, 1); dist = Vector{Float64}(undef, 1)
  GC.@preserve q idx dist check(tree.ctx, ccall((:rrtqx_nearest_batch, LIB), Int32,
      (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Int32}, Ptr{Float64}), tree.h, q, 1, idx, dist))
  return (tree.nodes[idx[1] + 1], dist[1])
end

function kdFindNearestBatch(tree::GpuKDTree, queries::Matrix{Float64})   # d x nq (column per query)
  nq = size(queries, 2); idx = Vector{Int32}(undef, nq); dist = Vector{Float64}(undef, nq)
  GC.@preserve queries idx dist check(tree.ctx, ccall((:rrtqx_nearest_batch, LIB), Int32,
      (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Int32}, Ptr{Float64}), tree.h, queries, nq, idx, dist))
  return (idx .+ Int32(1), dist)
end

# batched range query -> (counts, offsets (1-based starts), idx (1-based), dist)
function kdFindWithinRangeBatch(tree::GpuKDTree, range::Float64, queries::Matrix{Float64};
                                ranges::Union{Nothing,Vector{Float64}} = nothing, wantDist::Bool = true)
  nq = size(queries, 2); total = Ref{Int64}(0)
  flags = wantDist ? RANGE_WANT_DIST : UInt32(0)
  rp = ranges === nothing ? Ptr{Float64}(C_NULL) : pointer(ranges)
  GC.@preserve queries ranges check(tree.ctx, ccall((:rrtqx_range_query_batch, LIB), Int32,
      (Ptr{Cvoid}, Ptr{Float64}, Int64, Float64, Ptr{Float64}, UInt32, Ref{Ptr{Cvoid}}, Ref{Int64}),
      tree.h, queries, nq, range, rp, flags, tree.res, total))
  counts = Vector{Int32}(undef, nq); offsets = Vector{Int64}(undef, nq)
  idx = Vector{Int32}(undef, total[]); dist = Vector{Float64}(undef, wantDist ? total[] : 0)
  GC.@preserve counts offsets idx dist begin
    check(tree.ctx, ccall((:rrtqx_range_result_layout, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int64}),
                          tree.res[], counts, offsets))
    check(tree.ctx, ccall((:rrtqx_range_result_fetch, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Float64}),
                          tree.res[], idx, wantDist ? pointer(dist) : Ptr{Float64}(C_NULL)))
  end
  return (counts, offsets .+ 1, idx .+ Int32(1), dist)
end

# addToRangeList (kdTree_general.jl:765-771) is reused from the reference: it dedups through inHeap.
# kdFindMoreWithinRange (kdTree_general.jl:927-955)
function kdFindMoreWithinRange(tree::GpuKDTree, range::Float64, queryPoint::Array{Float64}, L)
  tree.treeSize == 0 && error("kdFindWithinRange on an empty tree")
  q = reshape(vec(Array{Float64}(queryPoint)), tree.d, 1)
  (counts, offsets, idx, dist) = kdFindWithinRangeBatch(tree, range, q)
  for k = 1:counts[1]
    Main.addToRangeList(L, tree.nodes[idx[offsets[1] + k - 1]], dist[offsets[1] + k - 1])
  end
  return L
end

# kdFindWithinRange (kdTree_general.jl:889-919)
function kdFindWithinRange(tree::GpuKDTree{T}, range::Float64, queryPoint::Array{Float64}) where {T}
  return kdFindMoreWithinRange(tree, range, queryPoint, Main.JList{T}())
end

# ---------------------------------------------------------------- obstacles
# device mirror of S.obstacles (List{SphereObstacle}), front-to-back order
mutable struct GpuObstacles
  ctx::Context
  h::Ptr{Cvoid}
  ids::IdDict{Any,Int32}          # obstacle -> device index
  function GpuObstacles(ctx::Context)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, ccall((:rrtqx_spheres_create, LIB), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), ctx.h, h))
    o = new(ctx, h[], IdDict{Any,Int32}())
    finalizer(x -> ccall((:rrtqx_spheres_destroy, LIB), Int32, (Ptr{Cvoid},), x.h), o)
    return o
  end
end

# active[i] = !(obstacleUnused || lifeSpan <= 0): the early-out of explicitEdgeCheck3D (DRRT_Q.jl:1777)
function syncObstacles!(G::GpuObstacles, S)
  n = S.obstacles.length
  centers = Matrix{Float64}(undef, 3, n); radii = Vector{Float64}(undef, n); active = Vector{UInt8}(undef, n)
  empty!(G.ids)
  item = S.obstacles.front
  for i = 1:n
    ob = item.data
    centers[:, i] = ob.position[1:3]
    radii[i] = ob.radius
    active[i] = (ob.obstacleUnused || ob.lifeSpan <= 0) ? 0x00 : 0x01
    G.ids[ob] = Int32(i - 1)
    item = item.child
  end
  GC.@preserve centers radii active check(G.ctx, ccall((:rrtqx_spheres_upload, LIB), Int32,
      (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}, Int64), G.h, centers, radii, active, n))
  return G
end

# explicitEdgeCheck(S, edge) (DRRT_Q.jl:1802-1826), SimpleEdge
function explicitEdgeCheck(G::GpuObstacles, S, edge, flags::UInt32 = UInt32(0))
  S.inWarmupTime && return false
  s = vec(Array{Float64}(edge.startNode.position))[1:3]; e = vec(Array{Float64}(edge.endNode.position))[1:3]
  out = Vector{UInt8}(undef, 1)
  GC.@preserve s e out check(G.ctx, ccall((:rrtqx_segment_check_batch, LIB), Int32,
      (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Float64, UInt32, Ptr{UInt8}),
      G.ctx.h, G.h, s, e, 1, S.robotRadius, flags, out))
  return out[1] != 0x00
end

# batched explicitEdgeCheck over edges between tree nodes given as 1-based index vectors
function explicitEdgeCheckBatch(G::GpuObstacles, S, tree::GpuKDTree, src::Vector{Int32}, dst::Vector{Int32},
                                flags::UInt32 = UInt32(0))
  out = Vector{UInt8}(undef, length(src))
  S.inWarmupTime && return fill!(out, 0x00)
  s0 = src .- Int32(1); d0 = dst .- Int32(1)
  GC.@preserve s0 d0 out check(G.ctx, ccall((:rrtqx_edge_check_batch, LIB), Int32,
      (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Int64, Float64, UInt32, Ptr{UInt8}),
      tree.h, G.h, s0, d0, length(src), S.robotRadius, flags, out))
  return out
end

function pointCheck(G::GpuObstacles, S, point::Array{Float64}, flags::UInt32)
  S.inWarmupTime && return (false, Inf)
  p = vec(Array{Float64}(point))[1:3]; out = Vector{UInt8}(undef, 1); cert = Vector{Float64}(undef, 1)
  GC.@preserve p out cert check(G.ctx, ccall((:rrtqx_node_check_batch, LIB), Int32,
      (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int64, Float64, UInt32, Ptr{UInt8}, Ptr{Float64}),
      G.ctx.h, G.h, p, 1, S.robotRadius, flags, out, cert))
  return (out[1] != 0x00, cert[1])
end
explicitPointCheck(G::GpuObstacles, S, point::Array{Float64}) = pointCheck(G, S, point, CHECK_QUICK_PASS)   # DRRT_Q.jl:1520
explicitPointCheck3D(G::GpuObstacles, S, point::Array{Float64}) = pointCheck(G, S, point, UInt32(0))        # DRRT_Q.jl:1558
explicitNodeCheck(G::GpuObstacles, S, node) = explicitPointCheck(G, S, node.position)                       # DRRT_Q.jl:1594
explicitNodeCheck3D(G::GpuObstacles, S, node) = explicitPointCheck3D(G, S, node.position)                   # DRRT_Q.jl:1595

# -------------------------------------------------------------- edges + sweeps
mutable struct GpuEdges
  tree::GpuKDTree
  h::Ptr{Cvoid}
  items::Vector{Any}              # edge id + 1 -> JListNode holding the edge
  res::Base.RefValue{Ptr{Cvoid}}
  function GpuEdges(tree::GpuKDTree)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(tree.ctx, ccall((:rrtqx_edges_create, LIB), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), tree.h, h))
    e = new(tree, h[], Vector{Any}(), Ref{Ptr{Cvoid}}(C_NULL))
    finalizer(e) do x
      x.res[] != C_NULL && ccall((:rrtqx_sweep_result_destroy, LIB), Int32, (Ptr{Cvoid},), x.res[])
      ccall((:rrtqx_edges_destroy, LIB), Int32, (Ptr{Cvoid},), x.h)
    end
    return e
  end
end

# mirrors every node's out-edges in the order of RRTNodeNeighborIterator (DRRT_Q.jl:2408-2431):
# InitialNeighborListOut, then rrtNeighborsOut; plus the parent edges
function syncEdges!(E::GpuEdges)
  tree = E.tree
  src = Int32[]; dst = Int32[]; empty!(E.items)
  parent = fill(Int32(-1), length(tree.nodes))
  for n in tree.nodes
    for lst in (n.InitialNeighborListOut, n.rrtNeighborsOut)
      item = lst.front
      for k = 1:lst.length
        push!(src, tree.index[n]); push!(dst, tree.index[item.data.endNode]); push!(E.items, item)
        item = item.child
      end
    end
    if n.rrtParentUsed
      parent[tree.index[n] + 1] = tree.index[n.rrtParentEdge.endNode]
    end
  end
  GC.@preserve src dst parent check(tree.ctx, ccall((:rrtqx_edges_upload, LIB), Int32,
      (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Int64, Ptr{Int32}, Int64), E.h, src, dst, length(src), parent, length(parent)))
  return E
end

function fetchSweep(E::GpuEdges)
  ne = Ref{Int64}(0); nn = Ref{Int64}(0)
  check(E.tree.ctx, ccall((:rrtqx_sweep_result_sizes, LIB), Int32,
                          (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ptr{Int64}, Ptr{Int64}), E.res[], ne, nn, C_NULL, C_NULL))
  edges = Vector{Int32}(undef, ne[]); nodes = Vector{Int32}(undef, nn[])
  GC.@preserve edges nodes check(E.tree.ctx, ccall((:rrtqx_sweep_result_fetch, LIB), Int32,
                                                   (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}), E.res[], edges, nodes))
  return (edges, nodes)
end

# findPointsInConflictWithObstacle (DRRT_Q.jl:3195-3215), Euclidean space without time/theta
function findPointsInConflictWithObstacle(S, KD::GpuKDTree, ob, root)
  (!S.spaceHasTime && !S.spaceHasTheta) || error("this type of obstacle not coded for this type of space")
  searchRange = S.robotRadius + S.delta + ob.radius
  return kdFindWithinRange(KD, searchRange, ob.position)
end

# addNewObstacle (DRRT_Q.jl:3220-3290): the GPU returns the blocked edge ids and the orphaned node ids,
# the reference's list surgery is applied here unchanged.
function addNewObstacle(G::GpuObstacles, E::GpuEdges, S, KD::GpuKDTree, Q, ob, root, fileCounter::Int, R)
  ob.obstacleUnused = false
  syncObstacles!(G, S)
  if ob.lifeSpan > 0
    ids = Int32[G.ids[ob]]
    GC.@preserve ids check(KD.ctx, ccall((:rrtqx_obstacle_add_sweep, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Int64, Float64, Float64, UInt32, Ref{Ptr{Cvoid}}),
        E.h, G.h, ids, 1, S.robotRadius, S.delta, UInt32(0), E.res))
    (blocked, orphans) = fetchSweep(E)
    for e in blocked
      E.items[e + 1].data.dist = Inf                                   # :3248-3249
    end
    for v in orphans                                                   # :3257-3270
      thisNode = KD.nodes[v + 1]
      Main.JlistRemove(thisNode.rrtParentEdge.endNode.SuccessorList, thisNode.successorListItemInParent)
      thisNode.rrtParentEdge.endNode = thisNode
      thisNode.rrtParentEdge.dist = Inf
      thisNode.rrtParentUsed = false
      Main.verifyInOSQueue(Q, thisNode)
    end
  end
  if R.robotEdgeUsed                                                   # :3287-3289
    one = GpuObstacles(G.ctx)
    c = Array{Float64}(ob.position[1:3]); r = Float64[ob.radius]; a = UInt8[(ob.lifeSpan > 0) ? 0x01 : 0x00]
    GC.@preserve c r a check(G.ctx, ccall((:rrtqx_spheres_upload, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}, Int64), one.h, c, r, a, 1))
    if explicitEdgeCheck(one, S, R.robotEdge)
      R.currentMoveInvalid = true
    end
  end
end

# removeObstacle (DRRT_Q.jl:3295-3362).  qxSemantics = true reproduces this fork (the obstacle is disabled
# before the loop, so nothing is ever restored); false gives the Otte generation (DRRT.jl:3202-3268).
function removeObstacle(G::GpuObstacles, E::GpuEdges, S, KD::GpuKDTree, Q, ob, root, hyberBallRad::Float64,
                        timeElapsed::Float64, moveGoal; qxSemantics::Bool = true)
  syncObstacles!(G, S)
  obId = G.ids[ob]
  ob.expired = true
  ob.obstacleUnused = true
  others = Int32[]
  item = S.obstacles.front
  for i = 1:S.obstacles.length
    o = item.data
    if o != ob && !o.obstacleUnused && o.lifeSpan > 0 && o.startTime <= timeElapsed <= (o.startTime + o.lifeSpan)
      push!(others, G.ids[o])
    end
    item = item.child
  end
  inf = UInt8[(it.data.dist == Inf) ? 0x01 : 0x00 for it in E.items]
  flags = qxSemantics ? SWEEP_REMOVED_INACTIVE : UInt32(0)
  GC.@preserve others inf check(KD.ctx, ccall((:rrtqx_obstacle_remove_sweep, LIB), Int32,
      (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Int32}, Int64, Ptr{UInt8}, Float64, Float64, UInt32, Ref{Ptr{Cvoid}}),
      E.h, G.h, obId, others, length(others), inf, S.robotRadius, S.delta, flags, E.res))
  (restored, requeue) = fetchSweep(E)
  for e in restored
    E.items[e + 1].data.dist = E.items[e + 1].data.distOriginal        # :3340-3346
  end
  for v in requeue                                                     # :3352-3357
    thisNode = KD.nodes[v + 1]
    Main.recalculateLMCMineVTwo(Q, thisNode, root, hyberBallRad)
    if thisNode.rrtTreeCost != thisNode.rrtLMC && Main.lessQ(thisNode, moveGoal)
      Main.verifyInQueue(Q, thisNode)
    end
  end
  ob.obstacleUnused = true
end

# ------------------------------------------------------------------ neighbour-graph residency
# Edges the planner creates in one iteration (makeNeighborOf / makeInitialOutNeighborOf,
# DRRT_Q.jl:2589-2593) are appended to the resident set instead of re-uploading the graph; parent
# edges are re-pointed in place (makeParentOf, DRRT_Q.jl:1841-1856).  `srcIdx`/`dstIdx` are 0-based
# device node indices (KD.index[node]); edge ids continue the order of E.items.
function appendEdges!(E::GpuEdges, newEdges::Vector, srcIdx::Vector{Int32}, dstIdx::Vector{Int32})
  append!(E.items, newEdges)
  GC.@preserve srcIdx dstIdx check(E.tree.ctx, ccall((:rrtqx_edges_append, LIB), Int32,
      (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Int64), E.h, srcIdx, dstIdx, length(srcIdx)))
end

function setParents!(E::GpuEdges, nodeIdx::Vector{Int32}, parentIdx::Vector{Int32})   # parent -1: rrtParentUsed = false
  GC.@preserve nodeIdx parentIdx check(E.tree.ctx, ccall((:rrtqx_edges_set_parents, LIB), Int32,
      (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Int64), E.h, nodeIdx, parentIdx, length(nodeIdx)))
end

# ------------------------------------------------------------------ fused per-iteration query
# One launch for what extend() asks of the geometry per sample (rrtqx.jl:926-950, DRRT_Q.jl:2551-2637):
# nearest node, explicitNodeCheck of the sample, the shrinking-ball neighbours with their keys, and
# explicitEdgeCheck of every new edge in both directions.
function extendQuery(G::GpuObstacles, S, tree::GpuKDTree, point::Array{Float64}, range::Float64; capacity::Int = 8192)
  nearestIdx = Ref{Int32}(0); nearestDist = Ref{Float64}(0.0)
  collides = Ref{UInt8}(0); cert = Ref{Float64}(0.0); n = Ref{Int32}(0)
  idx = Vector{Int32}(undef, capacity); dist = Vector{Float64}(undef, capacity)
  fwd = Vector{UInt8}(undef, capacity); rev = Vector{UInt8}(undef, capacity)
  p = vec(point)
  GC.@preserve p idx dist fwd rev check(tree.ctx, ccall((:rrtqx_extend_query, LIB), Int32,
      (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Float64, Float64, UInt32, Int32, Ref{Int32}, Ref{Float64}, Ref{UInt8},
       Ref{Float64}, Ref{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{UInt8}, Ptr{UInt8}),
      tree.h, G.h, p, range, S.robotRadius, UInt32(4), Int32(capacity), nearestIdx, nearestDist, collides, cert, n,
      idx, dist, fwd, rev))
  k = Int(n[])
  return (tree.nodes[nearestIdx[] + 1], nearestDist[], collides[] != 0, cert[],
          idx[1:k], dist[1:k], fwd[1:k] .!= 0, rev[1:k] .!= 0)
end

# ------------------------------------------------------------------ Dubins edges (2-D polygon world)
mutable struct GpuPolygons
  ctx::Context
  h::Ptr{Cvoid}
  function GpuPolygons(ctx::Context)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, ccall((:rrtqx_polygons_create, LIB), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), ctx.h, h))
    o = new(ctx, h[])
    finalizer(x -> ccall((:rrtqx_polygons_destroy, LIB), Int32, (Ptr{Cvoid},), x.h), o)
    return o
  end
end

# Mirrors S.obstacles (Obstacle kinds 1 and 3, DRRT_data_structures.jl:135-265): bounding circles as the
# constructor computed them (ob.position, ob.radius), polygon vertices as a CSR.
function syncPolygons!(G::GpuPolygons, S)
  kinds = Int32[]; centers = Float64[]; radii = Float64[]; active = UInt8[]; vptr = Int64[0]; verts = Float64[]
  item = S.obstacles.front
  for i = 1:S.obstacles.length
    ob = item.data
    push!(kinds, Int32(ob.kind)); push!(centers, ob.position[1], ob.position[2]); push!(radii, ob.radius)
    push!(active, (ob.obstacleUnused || ob.lifeSpan <= 0) ? 0x00 : 0x01)
    if ob.kind == 3
      for r = 1:size(ob.polygon, 1)
        push!(verts, ob.polygon[r, 1], ob.polygon[r, 2])
      end
    end
    push!(vptr, length(verts) ÷ 2)
    item = item.child
  end
  GC.@preserve kinds centers radii active vptr verts check(G.ctx, ccall((:rrtqx_polygons_upload, LIB), Int32,
      (Ptr{Cvoid}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}, Ptr{Int64}, Ptr{Float64}, Int64),
      G.h, kinds, centers, radii, active, vptr, verts, length(kinds)))
end

const DUBINS_TYPES = ("rsl", "rsr", "rlr", "lsr", "lsl", "lrl")

# calculateTrajectory(S, edge::DubinsEdge) for a batch of edges (DRRT_DubinsEdge_functions.jl:329-709, space
# without time): sets dist / distOriginal / Wdist / dubinsType / trajectory on every edge.
function calculateTrajectoryBatch(ctx::Context, S, edges::Vector)
  n = length(edges)
  starts = Matrix{Float64}(undef, 4, n); goals = Matrix{Float64}(undef, 4, n)
  for (i, e) in enumerate(edges)
    starts[:, i] = e.startNode.position[1:4]; goals[:, i] = e.endNode.position[1:4]
  end
  res = Ref{Ptr{Cvoid}}(C_NULL)
  GC.@preserve starts goals check(ctx, ccall((:rrtqx_dubins_trajectory_batch, LIB), Int32,
      (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Float64, Ref{Ptr{Cvoid}}), ctx.h, starts, goals, n, S.minTurningRadius, res))
  ne = Ref{Int64}(0); nr = Ref{Int64}(0)
  check(ctx, ccall((:rrtqx_dubins_result_sizes, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}), res[], ne, nr))
  dist = Vector{Float64}(undef, n); typ = Vector{Int32}(undef, n); ptr = Vector{Int64}(undef, n + 1)
  xy = Matrix{Float64}(undef, 2, nr[])
  GC.@preserve dist typ ptr xy check(ctx, ccall((:rrtqx_dubins_result_fetch, LIB), Int32,
      (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int32}, Ptr{Int64}, Ptr{Float64}), res[], dist, typ, ptr, xy))
  ccall((:rrtqx_dubins_result_destroy, LIB), Int32, (Ptr{Cvoid},), res[])
  for (i, e) in enumerate(edges)
    e.dubinsType = typ[i] < 0 ? "xxx" : DUBINS_TYPES[typ[i] + 1]
    e.Wdist = dist[i]; e.dist = dist[i]; e.distOriginal = dist[i]
    e.trajectory = permutedims(xy[:, ptr[i]+1:ptr[i+1]])          # rows x 2, as the reference builds it
  end
end

# Dubins explicitEdgeCheck(S, edge) OR-ed over the obstacle list for a batch (DRRT_DubinsEdge_functions.jl:750-774).
function explicitEdgeCheckBatch(G::GpuPolygons, S, edges::Vector)
  n = length(edges)
  starts = Matrix{Float64}(undef, 2, n); ends = Matrix{Float64}(undef, 2, n); ptr = Vector{Int64}(undef, n + 1); ptr[1] = 0
  for (i, e) in enumerate(edges)
    starts[:, i] = e.startNode.position[1:2]; ends[:, i] = e.endNode.position[1:2]
    ptr[i + 1] = ptr[i] + size(e.trajectory, 1)
  end
  xy = Matrix{Float64}(undef, 2, ptr[end])
  for (i, e) in enumerate(edges)
    xy[:, ptr[i]+1:ptr[i+1]] = permutedims(e.trajectory[:, 1:2])
  end
  out = Vector{UInt8}(undef, n)
  GC.@preserve starts ends ptr xy out check(G.ctx, ccall((:rrtqx_dubins_edge_check_batch, LIB), Int32,
      (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Int64, Float64, Float64, UInt32, Ptr{UInt8}),
      G.h, starts, ends, ptr, xy, n, S.robotRadius, S.minTurningRadius, UInt32(0), out))
  return S.inWarmupTime ? falses(n) : out .!= 0
end

# saturate(newPoint, closestPoint, delta), DubinsEdge version (DRRT_DubinsEdge_functions.jl:70-95): in place.
function saturateDubins!(ctx::Context, newPoint::Array{Float64}, closestPoint::Array{Float64}, delta::Float64)
  p = vec(newPoint); c = vec(closestPoint)
  GC.@preserve p c check(ctx, ccall((:rrtqx_dubins_saturate_batch, LIB), Int32,
      (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Float64), ctx.h, p, c, 1, delta))
  newPoint[:] = p
end

end # module
