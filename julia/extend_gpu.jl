# extend_gpu.jl -- the consumer of extendQuery: `extend` (rrtXQueue version, DRRT_Q.jl:2546-2641) and
# `findBestParent` (DRRT_Q.jl:1927-1990) rewritten over the arrays ONE fused launch returns, instead of one blocking
# explicitEdgeCheck launch per edge (2 k of them per iteration, k ~ 300 at 20 000 nodes).
#
# Included by RRTQXGpu.jl (inside module RRTQXGpu).  RRTQXGpu.install!() forwards
#     extend(S, KD::GpuKDTree, Q::rrtXQueue, newNode, closestNode, delta, hyberBallRad, moveGoal)
# to extendRRTx below, so the planner loop (rrtqx.jl:926-950) is unchanged.  Every mutation of the planner's
# structures is the reference's own line, in the reference's order; only the SOURCE of three values changes:
#     kdFindWithinRange(KD, hyberBallRad, newNode.position)     -> neighbour indices + keys of extendQuery
#     explicitEdgeCheck(S, edge(newNode -> nearNode))           -> fwd[i]
#     explicitEdgeCheck(S, edge(nearNode -> newNode))           -> rev[i]
# The neighbour ORDER is the device's, not the reference's LIFO traversal order; the planner is order-insensitive
# except under exact cost ties in findBestParent (DRRT_Q.jl:1970, SURVEY.md appendix B3).
#
# Written blind (no Julia in the build image), like RRTQXGpu.jl; SimpleEdge worlds (sphere obstacles).

# findBestParent(S, newNode, nodeList, closestNode, saveAllEdges = true) over the query arrays.
# nearNodes[i] is the i-th neighbour, fwd[i] = explicitEdgeCheck(S, edge(newNode -> nearNodes[i])).
function findBestParentGpu(S, newNode, nearNodes::Vector, fwd, closestNode)
  # update LMC value based on nodes in the list                         (DRRT_Q.jl:1936-1939)
  newNode.rrtLMC = Inf
  newNode.rrtTreeCost = Inf
  newNode.rrtParentUsed = false
  for i = 1:length(nearNodes)
    nearNode = nearNodes[i]
    thisEdge = Main.newEdge(newNode, nearNode)                          # :1949
    Main.calculateTrajectory(S, thisEdge)                               # :1950 (SimpleEdge: three distances)
    nearNode.tempEdge = thisEdge                                        # :1952-1954, saveAllEdges
    if fwd[i] || !Main.validMove(S, thisEdge)                           # :1959
      nearNode.tempEdge.dist = Inf                                      # :1961-1963
      continue
    end
    if newNode.rrtLMC > nearNode.rrtLMC + thisEdge.dist                 # :1970-1975
      newNode.rrtLMC = nearNode.rrtLMC + thisEdge.dist
      newNode.rrtParentEdge = thisEdge
      newNode.rrtParentUsed = true
    end
  end
end

# extend(S, KD, Q::rrtXQueue, newNode, closestNode, delta, hyberBallRad, moveGoal), DRRT_Q.jl:2546-2641
function extendRRTx(S, KD::GpuKDTree, Q, newNode, closestNode, delta::Float64, hyberBallRad::Float64, moveGoal)
  # one launch: the shrinking-ball neighbours of newNode with keys and both-direction collision flags
  (_, _, _, _, idx, keys, fwdv, revv) = extendQuery(S, KD, newNode.position, hyberBallRad)
  k = length(idx)
  nearNodes = Vector{Any}(undef, k)
  fwd = Vector{Bool}(undef, k); rev = Vector{Bool}(undef, k); key = Vector{Float64}(undef, k)
  for i = 1:k                       # copy out of the pinned views: the per-edge fall-backs below may query again
    nearNodes[i] = KD.nodes[idx[i] + 1]
    fwd[i] = fwdv[i]; rev[i] = revv[i]; key[i] = keys[i]
  end
  if k == 0 && S.goalNode != newNode                                    # DRRT_Q.jl:1930-1934: empty ball -> closestNode
    nearNodes = Any[closestNode]
    e1 = Main.newEdge(newNode, closestNode); e2 = Main.newEdge(closestNode, newNode)
    fwd = Bool[explicitEdgeCheck(S, e1)]; rev = Bool[explicitEdgeCheck(S, e2)]   # two per-edge launches, early planning only
    key = Float64[0.0]              # the reference leaves this key uninitialised (2-argument JlistPush, jlist.jl:56-77)
    k = 1
  end

  # try to find and link to best parent; saves the edges newNode -> neighbour in nearNode.tempEdge
  findBestParentGpu(S, newNode, nearNodes, fwd, closestNode)

  # if no parent was found then ignore this node                        (:2560-2563)
  newNode.rrtParentUsed || return

  # add the new node to its parent's successor list                     (:2569-2573)
  parentNode = newNode.rrtParentEdge.endNode
  backEdge = Main.newEdge(parentNode, newNode)
  backEdge.dist = Inf
  Main.JlistPush(parentNode.SuccessorList, backEdge, Inf)
  newNode.successorListItemInParent = parentNode.SuccessorList.front

  kdInsert(KD, newNode)                                                 # :2575 (one asynchronous launch)

  for i = 1:k                                                           # :2581-2636
    nearNode = nearNodes[i]
    if key[i] != Inf                # :2586 tests the RANGE KEY, never Inf: collided forward edges are linked too (B7)
      Main.makeInitialOutNeighborOf(nearNode, newNode, nearNode.tempEdge)
      Main.makeNeighborOf(nearNode, newNode, nearNode.tempEdge)
    end
    thisEdge = Main.newEdge(nearNode, newNode)                          # :2600-2601
    Main.calculateTrajectory(S, thisEdge)
    if Main.validMove(S, thisEdge) && !rev[i]                           # :2603
      Main.makeInitialInNeighborOf(newNode, nearNode, thisEdge)
      Main.makeNeighborOf(newNode, nearNode, thisEdge)
    else
      continue
    end
    if (nearNode.rrtLMC > newNode.rrtLMC + thisEdge.dist &&
        newNode.rrtParentEdge.endNode != nearNode &&
        newNode.rrtLMC + thisEdge.dist < moveGoal.rrtLMC)               # :2618-2620
      Main.makeParentOf(newNode, nearNode, thisEdge, KD.root)
      oldLmc = nearNode.rrtLMC
      nearNode.rrtLMC = newNode.rrtLMC + thisEdge.dist
      if oldLmc - nearNode.rrtLMC > Q.changeThresh && nearNode != KD.root
        Main.verifyInQueue(Q, nearNode)
      end
    end
  end
  # (emptyRangeList: nothing to clean up, no inHeap marks were set)
  Main.addToHeap(Q.Q, newNode)                                          # :2640
  markEdgesStale!(S)                # neighbour lists / parents changed: the resident edge set re-syncs before the next sweep
  return
end
