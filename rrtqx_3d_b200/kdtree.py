"""KDTree and the kd* functions with the reference's names and call shapes
(kdTree_general.jl), running on the GPU through the C ABI.

    tree = KDTree(ctx, d, KDdist)                      # KDTree{T}(d, f)
    tree = KDTree(ctx, 4, KDdist, [4], [2*pi])         # KDTree{T}(d, f, wraps, wrapPoints) (1-based dims)
    kdInsert(tree, node)
    (node, dist) = kdFindNearest(tree, queryPoint)
    L = kdFindWithinRange(tree, range, queryPoint)     # JList with key = dist
    kdFindMoreWithinRange(tree, range, queryPoint, L)
    (node, key) = popFromRangeList(L);  emptyRangeList(L)

Batched twins (`kdFindNearestBatch`, `kdFindWithinRangeBatch`) are what a planner
restructured around batches would call; the single-query functions are the drop-in
shims over the same kernels.
"""
from __future__ import annotations

import numpy as np

from .device import Context, DeviceTree, RangeResult
from .structures import JList, RRTNode


def euclidianDist(x, y) -> float:
    """DRRT_distance_functions.jl:37 (host helper; left-to-right sum of squares)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    s = (x[0] - y[0]) * (x[0] - y[0])
    for i in range(1, x.size):
        s = s + (x[i] - y[i]) * (x[i] - y[i])
    return float(np.sqrt(s))


KDdist = euclidianDist  # DRRT_SimpleEdge_functions.jl:51 / DRRT_DubinsEdge_functions.jl:52


class KDTree:
    """KDTree{T} (kdTree_general.jl:94-112) backed by a device-resident tree."""

    def __init__(self, ctx: Context, d: int, distanceFunction=KDdist, wraps=(), wrapPoints=()):
        if distanceFunction not in (KDdist, euclidianDist):
            raise ValueError("the GPU kd tree implements KDdist = euclidianDist only (both edge files use it)")
        self.ctx = ctx
        self.d = int(d)
        self.distanceFunction = distanceFunction
        self.treeSize = 0
        self.wraps = [int(w) for w in wraps]           # 1-based, as in Julia
        self.wrapPoints = [float(w) for w in wrapPoints]
        self.numWraps = len(self.wraps)
        self.root = None
        self.nodes = []                                 # device index -> node
        self.dev = DeviceTree(ctx, self.d, [w - 1 for w in self.wraps], self.wrapPoints)
        self._res = RangeResult(ctx)

    # -- internal ---------------------------------------------------------
    def _refresh_kd_fields(self, first: int, count: int):
        """Keep kdParent / kdChildL / kdChildR / kdSplit of the Julia-visible nodes populated
        (DRRT_data_structures.jl:25-33, 68-72) so traversals such as saveRRTTree keep working."""
        parent, cl, cr, split = self.dev.kd_fields(first, count)
        for k in range(count):
            n = self.nodes[first + k]
            n.kdSplit = int(split[k]) + 1               # Julia dims are 1-based
            if parent[k] >= 0:
                p = self.nodes[parent[k]]
                n.kdParent = p
                n.kdParentExist = True
                if p.position[0, p.kdSplit - 1] > n.position[0, p.kdSplit - 1]:
                    p.kdChildL, p.kdChildLExist = n, True
                else:
                    p.kdChildR, p.kdChildRExist = n, True


def kdInsert(tree: KDTree, node: RRTNode):
    """kdInsert (kdTree_general.jl:121-170): no-op if node.kdInTree."""
    if node.kdInTree:
        return
    node.kdInTree = True
    idx = tree.dev.insert(node.position)
    node.kdIndex = idx
    tree.nodes.append(node)
    if tree.treeSize == 0:
        tree.root = node
    tree.treeSize += 1
    tree._refresh_kd_fields(idx, 1)


def kdInsertBatch(tree: KDTree, nodes):
    """kdInsert for a sequence of nodes in order, one device call."""
    new = [n for n in nodes if not n.kdInTree]
    if not new:
        return
    pos = np.concatenate([n.position.reshape(1, -1) for n in new], axis=0)
    first = tree.dev.insert_batch(pos)
    for k, n in enumerate(new):
        n.kdInTree = True
        n.kdIndex = first + k
        tree.nodes.append(n)
    if tree.treeSize == 0:
        tree.root = new[0]
    tree.treeSize += len(new)
    tree._refresh_kd_fields(first, len(new))


def kdFindNearest(tree: KDTree, queryPoint):
    """kdFindNearest (kdTree_general.jl:357-385) -> (node, dist)."""
    q = np.ascontiguousarray(queryPoint, dtype=np.float64).reshape(1, tree.d)
    idx, dist = tree.dev.nearest(q)
    return tree.nodes[int(idx[0])], float(dist[0])


def kdFindNearestBatch(tree: KDTree, queryPoints):
    idx, dist = tree.dev.nearest(np.ascontiguousarray(queryPoints, dtype=np.float64).reshape(-1, tree.d))
    return [tree.nodes[int(i)] for i in idx], dist


def addToRangeList(S: JList, thisNode: RRTNode, key: float):
    """addToRangeList (kdTree_general.jl:765-771)."""
    if thisNode.inHeap:
        return
    thisNode.inHeap = True
    S.push(thisNode, key)


def popFromRangeList(S: JList):
    """popFromRangeList (kdTree_general.jl:774-779)."""
    thisNode, key = S.pop_key()
    thisNode.inHeap = False
    return thisNode, key


def emptyRangeList(S: JList):
    """emptyRangeList (kdTree_general.jl:782-787)."""
    while S.length > 0:
        thisNode, _ = S.pop_key()
        thisNode.inHeap = False


def kdFindMoreWithinRange(tree: KDTree, range_: float, queryPoint, L: JList) -> JList:
    """kdFindMoreWithinRange (kdTree_general.jl:927-955): nodes already in a range list
    (inHeap set) are skipped, exactly like the reference's addToRangeList dedup."""
    q = np.ascontiguousarray(queryPoint, dtype=np.float64).reshape(1, tree.d)
    res, _ = tree.dev.range_query(q, float(range_), want_dist=True, result=tree._res)
    (idx, dist), = res.lists()
    for i, k in zip(idx, dist):
        addToRangeList(L, tree.nodes[int(i)], float(k))
    return L


def kdFindWithinRange(tree: KDTree, range_: float, queryPoint) -> JList:
    """kdFindWithinRange (kdTree_general.jl:889-919) -> JList{node} with key = dist.
    The members carry inHeap = true until the list is emptied (reference behaviour)."""
    return kdFindMoreWithinRange(tree, range_, queryPoint, JList())


def kdFindWithinRangeBatch(tree: KDTree, range_, queryPoints, ranges=None, want_dist=True):
    """Batched form: returns (counts, offsets, idx, dist) numpy arrays (device indices)."""
    q = np.ascontiguousarray(queryPoints, dtype=np.float64).reshape(-1, tree.d)
    res, _ = tree.dev.range_query(q, float(range_), ranges=ranges, want_dist=want_dist, result=tree._res)
    counts, offsets = res.layout()
    idx, dist = res.fetch(want_dist)
    return counts, offsets, idx, dist
