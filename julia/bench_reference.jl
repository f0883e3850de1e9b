# bench_reference.jl -- times the reference's OWN kdFindWithinRange (kdTree_general.jl:889) on a sample of the C2
# workload; used by `bench.py --impl reference` when a `julia` binary and RRTQX_REFERENCE_DIR are present.
#   julia julia/bench_reference.jl <code_RRTQx_3D> <points.f64> <queries.f64> <n_points> <n_queries> <radius> <seconds>
# points / queries: raw little-endian Float64, row-major n x 3.  Prints one line: queries_per_s=<v> queries=<n> neighbours=<k>
# Never executed in the image this repository was built in (no Julia there); single-threaded like the reference.
const REF = ARGS[1]
for f in ("heap.jl", "list.jl", "jlist.jl", "kdTree_general.jl", "DRRT_distance_functions.jl", "DRRT_SimpleEdge.jl")
  include(joinpath(REF, f))
end
const Edge{T} = SimpleEdge{T}
for f in ("DRRT_data_structures.jl", "DRRT_SimpleEdge_functions.jl")
  include(joinpath(REF, f))
end

function main()
  np = parse(Int, ARGS[4]); nq = parse(Int, ARGS[5]); r = parse(Float64, ARGS[6]); budget = parse(Float64, ARGS[7])
  pts = Matrix{Float64}(undef, 3, np); read!(ARGS[2], pts)
  qs = Matrix{Float64}(undef, 3, nq); read!(ARGS[3], qs)
  KD = KDTree{RRTNode{Float64}}(3, KDdist)
  for i = 1:np
    kdInsert(KD, RRTNode{Float64}(reshape(pts[:, i], 1, 3)))
  end
  done = 0; total = 0
  L = kdFindWithinRange(KD, r, reshape(qs[:, 1], 1, 3)); emptyRangeList(L)      # compile
  t0 = time()
  while done < nq && time() - t0 < budget
    L = kdFindWithinRange(KD, r, reshape(qs[:, done + 1], 1, 3))
    total += L.length
    emptyRangeList(L)
    done += 1
  end
  dt = time() - t0
  println("queries_per_s=", done / dt, " queries=", done, " neighbours=", total)
end
main()
