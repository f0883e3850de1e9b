"""Host-side data structures with the reference's names, fields and semantics.

These are the Python twins of the Julia structs the planner keeps
(jlist.jl:30-54, DRRT_data_structures.jl:22-103, 267-307, 314-398,
DRRT_SimpleEdge.jl:34-56).  They are deliberately logic-free: the GPU path reads
`position`, the flags and the edge end points from them and writes results back
through the same fields the reference mutates.
"""
from __future__ import annotations

import math
from typing import Any, Iterator, Optional

import numpy as np


# ---------------------------------------------------------------- JList (jlist.jl)
class JListNode:
    __slots__ = ("child", "parent", "data", "key")

    def __init__(self):
        self.child = self      # jlist.jl:46-48: the bound node points to itself
        self.parent = self
        self.data = None
        self.key = 0.0


class JList:
    """Doubly linked list with keys; push/pop at the front (jlist.jl:41-160)."""

    def __init__(self):
        end = JListNode()
        self.front = end
        self.back = end
        self.bound = end
        self.length = 0

    def push(self, data: Any, key: Optional[float] = None) -> JListNode:
        """JlistPush (jlist.jl:56-97)."""
        n = JListNode()
        n.parent = self.front.parent
        n.child = self.front
        if self.length == 0:
            self.back = n
        else:
            self.front.parent = n
        n.data = data
        if key is not None:
            n.key = key
        self.front = n
        self.length += 1
        return n

    def top(self):
        """JlistTop (jlist.jl:99-105): false on empty."""
        return False if self.length == 0 else self.front.data

    def pop_key(self):
        """JlistPopKey (jlist.jl:138-160) -> (data, key); false on empty (jlist.jl:115-118)."""
        if self.length == 0:
            return False
        old = self.front
        if self.length > 1:
            self.front.child.parent = self.front.parent
            self.front = self.front.child
        elif self.length == 1:
            self.back = self.bound
            self.front = self.bound
        self.length -= 1
        old.child = old    # "in case Jlist nodes hang around after this" (jlist.jl:155-156)
        old.parent = old
        return old.data, old.key

    def pop(self):
        r = self.pop_key()
        return False if r is False else r[0]

    def remove(self, node: JListNode) -> bool:
        """JlistRemove (jlist.jl:162-190): unlink an arbitrary list node."""
        if self.length == 0:
            return True
        if self.front is node:
            self.front = node.child
        if self.back is node:
            self.back = node.parent
        nxt, prev = node.child, node.parent
        if self.length > 1 and prev is not prev.child:
            prev.child = nxt
        if self.length > 1 and nxt is not nxt.parent:
            nxt.parent = prev
        self.length -= 1
        if self.length == 0:
            self.back = self.bound
            self.front = self.bound
        node.parent = node
        node.child = node
        return True

    def __len__(self):
        return self.length

    def __iter__(self) -> Iterator[JListNode]:
        p = self.front
        for _ in range(self.length):
            yield p
            p = p.child

    def items(self):
        return [(n.data, n.key) for n in self]


# ------------------------------------------------------- edges (DRRT_SimpleEdge.jl)
class SimpleEdge:
    """SimpleEdge{T} (DRRT_SimpleEdge.jl:34-56)."""
    __slots__ = ("startNode", "endNode", "dist", "distOriginal", "Wdist", "listItemInStartNode",
                 "listItemInEndNode", "edgeId")

    def __init__(self, startNode=None, endNode=None):
        self.startNode = startNode
        self.endNode = endNode
        self.dist = 0.0
        self.distOriginal = 0.0
        self.Wdist = 0.0
        self.listItemInStartNode = None
        self.listItemInEndNode = None
        self.edgeId = -1  # position in the device edge set, when mirrored


Edge = SimpleEdge  # the reference aliases Edge{T} = SimpleEdge{T} (experimentsForRRTQX.jl:2-16)


def newEdge(startNode, endNode) -> SimpleEdge:
    """newEdge (DRRT_SimpleEdge_functions.jl:82-87)."""
    return Edge(startNode, endNode)


# ------------------------------------------------ nodes (DRRT_data_structures.jl:22)
class RRTNode:
    """RRTNode{T}: kd fields + graph lists + costs (DRRT_data_structures.jl:22-103)."""

    def __init__(self, position=None):
        self.kdInTree = False
        self.kdParentExist = False
        self.kdChildLExist = False
        self.kdChildRExist = False
        self.heapIndex = -1
        self.inHeap = False
        self.rrtParentUsed = False
        self.rrtNeighborsOut = JList()
        self.rrtNeighborsIn = JList()
        self.priorityQueueIndex = -1
        self.inPriorityQueue = False
        self.SuccessorList = JList()
        self.InitialNeighborListOut = JList()
        self.InitialNeighborListIn = JList()
        self.inOSQueue = False
        self.isMoveGoal = False
        self.position = None if position is None else np.ascontiguousarray(position, dtype=np.float64).reshape(1, -1)
        self.kdSplit = 0
        self.kdParent = None
        self.kdChildL = None
        self.kdChildR = None
        self.rrtParentEdge = None
        self.rrtTreeCost = math.inf
        self.rrtLMC = math.inf
        self.rrtH = 0.0
        self.tempEdge = None
        self.successorListItemInParent = None
        self.kdIndex = -1  # 0-based device index (the Julia module keeps it in an IdDict)


# ---------------------------------------- obstacles (DRRT_data_structures.jl:267-307)
class SphereObstacle:
    """SphereObstacle (DRRT_data_structures.jl:267-307)."""

    def __init__(self, position, radius: float):
        self.startTime = 0.0
        self.lifeSpan = math.inf
        self.obstacleUnused = False
        self.expired = False
        self.senseableObstacle = False
        self.obstacleUnusedAfterSense = True
        self.position = np.ascontiguousarray(position, dtype=np.float64).reshape(1, 3)
        self.radius = float(radius)
        self.radiusWithoutAug = float(radius)
        self.deviceId = -1  # position in the device sphere set

    def active(self) -> bool:
        """The early-out predicate of explicitEdgeCheck3D / explicitPointCheck2D
        (DRRT_Q.jl:1777, 1467): checked iff !(obstacleUnused || lifeSpan <= 0)."""
        return not (self.obstacleUnused or self.lifeSpan <= 0)


class ObstacleList:
    """List{SphereObstacle} (list.jl:41-59): push at the front, iterate front to back."""

    def __init__(self):
        self.items = []  # items[0] is the front

    def push(self, ob):
        self.items.insert(0, ob)

    @property
    def length(self):
        return len(self.items)

    def __iter__(self):
        return iter(self.items)

    def __len__(self):
        return len(self.items)


class CSpace:
    """CSpace{T} (DRRT_data_structures.jl:314-398), the fields the hot path reads."""

    def __init__(self, d: int, obsDelta, lowerBounds, upperBounds, start, goal):
        self.d = int(d)
        self.obstacles = ObstacleList()
        self.obsDelta = obsDelta
        self.lowerBounds = np.asarray(lowerBounds, dtype=np.float64).reshape(1, -1)
        self.upperBounds = np.asarray(upperBounds, dtype=np.float64).reshape(1, -1)
        self.width = self.upperBounds - self.lowerBounds
        self.start = np.asarray(start, dtype=np.float64).reshape(1, -1)
        self.goal = np.asarray(goal, dtype=np.float64).reshape(1, -1)
        self.spaceHasTime = False
        self.spaceHasTheta = False
        self.robotRadius = 0.0
        self.robotVelocity = 0.0
        self.delta = 0.0
        self.minTurningRadius = 0.0
        self.warmupTime = 0.0
        self.inWarmupTime = False
        self.pGoal = 0.0
        self.hypervolume = 0.0


def addObsToCSpace(C: CSpace, ob: SphereObstacle):
    """addObsToCSpace (DRRT_Q.jl:761): the obstacle list only grows, newest first."""
    C.obstacles.push(ob)


class RobotData:
    """RobotData fields the obstacle sweep reads (DRRT_data_structures.jl:456-535)."""

    def __init__(self):
        self.robotEdgeUsed = False
        self.robotEdge = None
        self.currentMoveInvalid = False
