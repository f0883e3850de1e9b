# make_reference_vectors.jl -- pins the CPU oracle of this repository against the REAL reference.
#
# Replays the seeded inputs of tests/golden/reference_inputs.txt (written by tests/golden/make_reference_inputs.py;
# no random numbers are drawn here) through the reference's own functions and writes their outputs, floats as IEEE
# bit patterns, to tests/golden/reference_vectors_<mode>.txt.  tests/test_reference_vectors.py then compares the
# oracle with those files (bit for bit; Dubins trajectories at 1e-9, SURVEY.md appendix A14) -- from that moment on
# parity is pinned to the reference instead of to a reading of its source.
#
# The reference cannot hold SimpleEdge and DubinsEdge in one session (`const Edge{T}` is set once), so there are two
# modes, each including the files exactly as the reference's experiment scripts do:
#
#     julia julia/make_reference_vectors.jl simple /path/to/RRTQX_3D/code_RRTQx_3D     # experimentsForRRTQX.jl:2-16
#     julia julia/make_reference_vectors.jl dubins /path/to/RRTQX_3D/code_RRTQx_3D     # dubinsExperimentsForPaper.jl:31-40
#
# Julia 1.0.5 is what the reference was written for; any 1.x should do (only Base + LinearAlgebra are used).
# This file has never been executed: there is no Julia in the image this repository was built in.  It is written
# against Julia 1.0 syntax and the reference's signatures, which are cited at every call.

using LinearAlgebra
using Printf

const MODE = length(ARGS) >= 1 ? ARGS[1] : "simple"
const REF = length(ARGS) >= 2 ? ARGS[2] : joinpath(@__DIR__, "..", "..", "reference", "code_RRTQx_3D")
const GOLDEN = joinpath(@__DIR__, "..", "tests", "golden")

# ------------------------------------------------------------------ container (tests/golden/refvec_io.py)
function read_arrays(path)
  out = Dict{String,Any}()
  lines = readlines(path)
  i = 1
  while i <= length(lines)
    ln = strip(lines[i]); i += 1
    (length(ln) > 0 && ln[1] == '@') || continue
    tok = split(ln[2:lastindex(ln)])
    name = String(tok[1]); rows = parse(Int, tok[2]); cols = parse(Int, tok[3]); kind = tok[4]
    if kind == "f"
      a = Matrix{Float64}(undef, rows, cols)
      for r = 1:rows
        t = split(lines[i + r - 1])
        for c = 1:cols
          a[r, c] = reinterpret(Float64, parse(UInt64, t[c], base = 16))
        end
      end
      out[name] = a
    else
      a = Matrix{Int64}(undef, rows, cols)
      for r = 1:rows
        t = split(lines[i + r - 1])
        for c = 1:cols
          a[r, c] = parse(Int64, t[c])
        end
      end
      out[name] = a
    end
    i += rows
  end
  return out
end

function write_array(io, name, a)
  if isa(a, AbstractVector)
    a = reshape(a, length(a), 1)
  end
  isfloat = eltype(a) <: AbstractFloat
  @printf(io, "@%s %d %d %s\n", name, size(a, 1), size(a, 2), isfloat ? "f" : "i")
  for r = 1:size(a, 1)
    if isfloat
      println(io, join([string(reinterpret(UInt64, Float64(a[r, c])), base = 16, pad = 16) for c = 1:size(a, 2)], " "))
    else
      println(io, join([string(Int64(a[r, c])) for c = 1:size(a, 2)], " "))
    end
  end
end

# ------------------------------------------------------------------ the reference, included as its scripts do
if MODE == "simple"
  for f in ("heap.jl", "list.jl", "jlist.jl", "kdTree_general.jl", "DRRT_distance_functions.jl", "DRRT_SimpleEdge.jl")
    include(joinpath(REF, f))
  end
  const Edge{T} = SimpleEdge{T}
  for f in ("DRRT_data_structures.jl", "DRRT_SimpleEdge_functions.jl", "DRRT_Q.jl")
    include(joinpath(REF, f))
  end
else
  for f in ("heap.jl", "list.jl", "jlist.jl", "kdTree_general.jl", "DRRT_distance_functions.jl", "DRRT_DubinsEdge.jl")
    include(joinpath(REF, f))
  end
  const Edge{T} = DubinsEdge{T}
  for f in ("DRRT_data_structures.jl", "DRRT_DubinsEdge_functions.jl", "DRRT.jl")
    include(joinpath(REF, f))
  end
end

const IN = read_arrays(joinpath(GOLDEN, "reference_inputs.txt"))
row(a, i) = reshape(a[i, :], 1, size(a, 2))        # positions are 1 x d matrices (DRRT_Q.jl:600)

# kd tree section: kdInsert / kdFindNearest / kdFindWithinRange (kdTree_general.jl:121,357,889), node ids = insert order
function kd_section(io, tag, pts, qs, radii, wraps, wrapPoints)
  d = size(pts, 2)
  KD = length(wraps) == 0 ? KDTree{RRTNode{Float64}}(d, KDdist) : KDTree{RRTNode{Float64}}(d, KDdist, wraps, wrapPoints)
  nodes = [RRTNode{Float64}(row(pts, i)) for i = 1:size(pts, 1)]
  id = IdDict{Any,Int}()
  for (i, n) in enumerate(nodes)
    kdInsert(KD, n); id[n] = i - 1
  end
  # kd topology of every node: parent / children (0-based, -1 absent), split (0-based)
  topo = Matrix{Int64}(undef, length(nodes), 4)
  for (i, n) in enumerate(nodes)
    topo[i, 1] = n.kdParentExist ? id[n.kdParent] : -1
    topo[i, 2] = n.kdChildLExist ? id[n.kdChildL] : -1
    topo[i, 3] = n.kdChildRExist ? id[n.kdChildR] : -1
    topo[i, 4] = n.kdSplit - 1
  end
  write_array(io, "$(tag)_topology", topo)
  nq = size(qs, 1)
  nnIdx = Vector{Int64}(undef, nq); nnDist = Vector{Float64}(undef, nq)
  for q = 1:nq
    (n, dd) = kdFindNearest(KD, row(qs, q))
    nnIdx[q] = id[n]; nnDist[q] = dd
  end
  write_array(io, "$(tag)_nn_idx", nnIdx); write_array(io, "$(tag)_nn_dist", nnDist)
  for (k, r) in enumerate(radii)
    counts = Vector{Int64}(undef, nq); idx = Int64[]; key = Float64[]
    for q = 1:nq
      L = kdFindWithinRange(KD, r, row(qs, q))
      counts[q] = L.length
      while L.length > 0                       # pop order (LIFO); the test compares sets + keys
        (n, kk) = popFromRangeList(L)
        push!(idx, id[n]); push!(key, kk)
      end
    end
    write_array(io, "$(tag)_range$(k)_count", counts)
    write_array(io, "$(tag)_range$(k)_idx", idx); write_array(io, "$(tag)_range$(k)_key", key)
  end
end

function run_simple(io)
  for d in (2, 3, 4)
    kd_section(io, "kd$(d)", IN["kd$(d)_pts"], IN["kd$(d)_qs"], vec(IN["kd_radii"]), Int[], Float64[])
  end
  kd_section(io, "kdw", IN["kdw_pts"], IN["kdw_qs"], vec(IN["kdw_radii"]), [4], [2.0 * pi])

  # ---- sphere world (DRRT_Q.jl:1205-1210, 1775-1826, 1520-1590)
  sph = IN["sph"]; unused = vec(IN["sph_unused"]); rho = IN["robot_radius"][1]
  S = CSpace{Float64}(3, -1.0, [-20.0 -20.0 -20.0], [20.0 20.0 20.0], [0.0 0.0 0.0], [1.0 1.0 1.0])
  S.robotRadius = rho
  S.spaceHasTime = false; S.spaceHasTheta = false
  obs = SphereObstacle[]
  for i = 1:size(sph, 1)
    ob = SphereObstacle(reshape(sph[i, 1:3], 1, 3), sph[i, 4])
    ob.obstacleUnused = unused[i] != 0
    push!(obs, ob)
  end
  for i = length(obs):-1:1                   # listPush puts the newest in front: list order = input order
    listPush(S.obstacles, obs[i])
  end
  ss = IN["seg_s"]; se = IN["seg_e"]; n = size(ss, 1)
  d2s = Vector{Float64}(undef, n); each = Matrix{Int64}(undef, n, length(obs)); all_ = Vector{Int64}(undef, n)
  for i = 1:n
    a = RRTNode{Float64}(row(ss, i)); b = RRTNode{Float64}(row(se, i))
    e = newEdge(a, b)                                                   # DRRT_SimpleEdge_functions.jl:82
    d2s[i] = distancePointToSegment(obs[1].position, a.position, b.position)   # DRRT_Q.jl:1205
    for (k, ob) in enumerate(obs)
      each[i, k] = explicitEdgeCheck3D(ob, a.position, b.position, rho) ? 1 : 0     # DRRT_Q.jl:1775
    end
    all_[i] = explicitEdgeCheck(S, e) ? 1 : 0                           # DRRT_Q.jl:1802
  end
  write_array(io, "seg_d2s", d2s); write_array(io, "seg_each", each); write_array(io, "seg_all", all_)
  P = IN["points"]; np_ = size(P, 1)
  pc = Matrix{Int64}(undef, np_, 2); cert = Matrix{Float64}(undef, np_, 2)
  for i = 1:np_
    (h, c) = explicitPointCheck(S, row(P, i));   pc[i, 1] = h ? 1 : 0; cert[i, 1] = c    # DRRT_Q.jl:1520
    (h, c) = explicitPointCheck3D(S, row(P, i)); pc[i, 2] = h ? 1 : 0; cert[i, 2] = c    # DRRT_Q.jl:1558
  end
  write_array(io, "point_hit", pc); write_array(io, "point_cert", cert)

  # ---- obstacle add / remove sweep (DRRT_Q.jl:3195-3362) on the 2000-node graph
  pts = IN["sw_pts"]; E = IN["sw_edges"]; par = vec(IN["sw_parent"]); S.delta = IN["sw_delta"][1]
  KD = KDTree{RRTNode{Float64}}(3, KDdist)
  nodes = [RRTNode{Float64}(row(pts, i)) for i = 1:size(pts, 1)]
  id = IdDict{Any,Int}()
  for (i, nd) in enumerate(nodes)
    kdInsert(KD, nd); id[nd] = i - 1
  end
  edges = Vector{Any}(undef, size(E, 1))
  for k = 1:size(E, 1)
    a = nodes[E[k, 1] + 1]; b = nodes[E[k, 2] + 1]
    e = newEdge(a, b); calculateTrajectory(S, e)                        # DRRT_SimpleEdge_functions.jl:177
    makeInitialOutNeighborOf(b, a, e)                                   # DRRT_Q.jl:2169: edge a -> b in a's out list
    edges[k] = e
  end
  Q = rrtXQueue{RRTNode{Float64}, typeof((Float64, Float64, Float64))}()
  Q.Q = BinaryHeap{RRTNode{Float64}, typeof((Float64, Float64, Float64))}(keyQ, lessQ, greaterQ, markQ, unmarkQ, markedQ, setIndexQ, unsetIndexQ, getIndexQ)
  Q.OS = JList{RRTNode{Float64}}(); Q.S = S; Q.changeThresh = 1.0
  R = RobotData{RRTNode{Float64}}([0.0 0.0 0.0], nodes[1], 10)
  so = IN["sw_obstacles"]
  for o = 1:size(so, 1)
    for (i, nd) in enumerate(nodes)          # fresh parents and edge costs for every obstacle
      nd.inOSQueue = false
      if par[i] >= 0
        pe = newEdge(nd, nodes[par[i] + 1]); calculateTrajectory(S, pe)
        nd.rrtParentEdge = pe; nd.rrtParentUsed = true
        JlistPush(nodes[par[i] + 1].SuccessorList, newEdge(nodes[par[i] + 1], nd), Inf)
        nd.successorListItemInParent = nodes[par[i] + 1].SuccessorList.front
      end
    end
    for e in edges
      e.dist = e.distOriginal
    end
    ob = SphereObstacle(reshape(so[o, 1:3], 1, 3), so[o, 4])
    ob.obstacleUnused = true                   # an appearing obstacle (DRRT_Q.jl:923-927)
    L = findPointsInConflictWithObstacle(S, KD, ob, nodes[1])           # DRRT_Q.jl:3195
    cand = Int64[]
    while L.length > 0
      (nd, kk) = popFromRangeList(L); push!(cand, id[nd])
    end
    addNewObstacle(S, KD, Q, ob, nodes[1], 0, R)                        # DRRT_Q.jl:3220
    blocked = Int64[k - 1 for k = 1:length(edges) if edges[k].dist == Inf]
    orphans = Int64[i - 1 for (i, nd) in enumerate(nodes) if par[i] >= 0 && !nd.rrtParentUsed]
    write_array(io, "sw$(o)_candidates", sort(cand)); write_array(io, "sw$(o)_blocked", blocked)
    write_array(io, "sw$(o)_orphans", orphans)
    listPush(S.obstacles, ob)
    removeObstacle(S, KD, Q, ob, nodes[1], 3.0, 0.0, nodes[1])          # DRRT_Q.jl:3295 (QX: nothing is restored)
    still = Int64[k - 1 for k = 1:length(edges) if edges[k].dist == Inf]
    write_array(io, "sw$(o)_blocked_after_remove", still)
  end
end

function run_dubins(io)
  # ---- 2-D polygon world (DRRT.jl:1060-1083, 1144-1202, 1523-1578; constructor DRRT_data_structures.jl:229-241)
  pp = vec(IN["poly_ptr"]); xy = IN["poly_xy"]; balls = IN["balls2d"]
  obs = Obstacle[]
  for i = 1:length(pp) - 1
    push!(obs, Obstacle(3, xy[pp[i]+1:pp[i+1], :]))
  end
  for i = 1:size(balls, 1)
    push!(obs, Obstacle(1, reshape(balls[i, 1:2], 1, 2), balls[i, 3]))
  end
  bound = Matrix{Float64}(undef, length(obs), 3)
  for (i, ob) in enumerate(obs)
    bound[i, 1] = ob.position[1]; bound[i, 2] = ob.position[2]; bound[i, 3] = ob.radius
  end
  write_array(io, "poly_bound", bound)
  ss = IN["seg2_s"]; se = IN["seg2_e"]; n = size(ss, 1)
  each = Matrix{Int64}(undef, n, length(obs)); pd = Vector{Float64}(undef, n); sd = Vector{Float64}(undef, n)
  A0 = vec(xy[1, :]); B0 = vec(xy[2, :])
  for i = 1:n
    s = row(ss, i); e = row(se, i)
    pd[i] = distanceSqrdPointToSegment(obs[1].position, vec(s), vec(e))              # DRRT.jl:1060
    sd[i] = segmentDistSqrd(vec(s), vec(e), A0, B0)                                   # DRRT.jl:1144
    for (k, ob) in enumerate(obs)
      each[i, k] = explicitEdgeCheck2D(ob, s, e, 0.5) ? 1 : 0                          # DRRT.jl:1523
    end
  end
  write_array(io, "seg2_pointdist", pd); write_array(io, "seg2_segdist", sd); write_array(io, "seg2_each", each)

  # ---- Dubins solver, trajectory check, saturate (DRRT_DubinsEdge_functions.jl:329-709, 750-774, 70-95)
  st = IN["dub_start"]; gl = IN["dub_goal"]; m = size(st, 1)
  S = CSpace{Float64}(4, -1.0, [-50.0 -50.0 0.0 0.0], [50.0 50.0 0.0 2.0 * pi], [0.0 0.0 0.0 0.0], [1.0 1.0 0.0 0.0])
  S.robotRadius = 0.5; S.minTurningRadius = IN["dub_rmin"][1]; S.spaceHasTime = false; S.spaceHasTheta = true
  S.robotVelocity = 1.0
  dist_ = Vector{Float64}(undef, m); typ = Vector{Int64}(undef, m); ptr = Int64[0]; traj = Vector{Any}()
  hit = Matrix{Int64}(undef, m, length(obs))
  names = Dict("rsl" => 0, "rsr" => 1, "rlr" => 2, "lsr" => 3, "lsl" => 4, "lrl" => 5)
  for i = 1:m
    e = newEdge(RRTNode{Float64}(row(st, i)), RRTNode{Float64}(row(gl, i)))
    calculateTrajectory(S, e)                                                         # :329
    dist_[i] = e.dist; typ[i] = get(names, e.dubinsType, -1)
    push!(traj, e.trajectory[:, 1:2]); push!(ptr, ptr[length(ptr)] + size(e.trajectory, 1))
    for (k, ob) in enumerate(obs)
      hit[i, k] = explicitEdgeCheck(S, e, ob) ? 1 : 0                                  # :750
    end
  end
  write_array(io, "dub_dist", dist_); write_array(io, "dub_type", typ); write_array(io, "dub_ptr", ptr)
  write_array(io, "dub_traj", vcat(traj...)); write_array(io, "dub_hit", hit)
  sat = Matrix{Float64}(undef, m, 4)
  for i = 1:m
    p = copy(row(st, i)); saturate(p, row(gl, i), IN["sat_delta"][1])                  # :70
    sat[i, :] = p
  end
  write_array(io, "dub_saturate", sat)
end

open(joinpath(GOLDEN, "reference_vectors_$(MODE).txt"), "w") do io
  println(io, "# written by julia/make_reference_vectors.jl $(MODE) with Julia $(VERSION)")
  MODE == "simple" ? run_simple(io) : run_dubins(io)
end
println("wrote ", joinpath(GOLDEN, "reference_vectors_$(MODE).txt"))
