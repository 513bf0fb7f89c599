"""TEST SCAFFOLDING: a small-n Python restatement of the GPU BVH builder's algorithm (luxcore_b200/csrc/build_kernels.cuh)
-- Morton order, Karras' radix tree or PLOC, greedy k-ary collapse, subtree sizes, the array index of a node as a sum
over its ancestors, emission -- so that the CPU suite can check the two things the kernels rely on without a GPU: that
the index formula puts every node exactly where a depth-first walk puts it, and that the array obeys the reference's
rules (bvhclassicbuild.cpp:181-220).  Pure Python loops: a few thousand leaves at most."""
import numpy as np

NODE_DTYPE = np.dtype([("w", "<u4", 6), ("nodeData", "<u4"), ("pad0", "<i4")])


def morton_order(boxes):
    c = (boxes[:, :3] + boxes[:, 3:]).astype(np.float64)
    lo, hi = c.min(0), c.max(0)
    ext = np.where(hi > lo, hi - lo, 1.0)
    q = np.minimum(2097151, ((c - lo) / ext * 2097152.0).astype(np.int64))
    code = np.zeros(len(boxes), dtype=object)
    for i in range(len(boxes)):
        v = 0
        for b in range(21):
            for a in range(3):
                if (int(q[i, a]) >> b) & 1:
                    v |= 1 << (3 * b + a)
        code[i] = v
    order = sorted(range(len(boxes)), key=lambda i: code[i])       # stable
    return np.asarray(order), [code[i] for i in order]


def radix_tree(keys):
    """Karras 2012.  ids: leaves 0 .. n-1 (sorted positions), inner n .. 2n-2, root = n."""
    n = len(keys)
    left = [-1] * (2 * n - 1); right = [-1] * (2 * n - 1); parent = [-1] * (2 * n - 1)

    def delta(i, j):
        if j < 0 or j >= n:
            return -1
        if keys[i] == keys[j]:
            return 64 + (32 - int(i ^ j).bit_length())
        return 64 - int(keys[i] ^ keys[j]).bit_length()
    for i in range(n - 1):
        d = 1 if delta(i, i + 1) - delta(i, i - 1) >= 0 else -1
        dmin = delta(i, i - d)
        lmax = 2
        while delta(i, i + lmax * d) > dmin:
            lmax <<= 1
        l, t = 0, lmax >> 1
        while t >= 1:
            if delta(i, i + (l + t) * d) > dmin:
                l += t
            t >>= 1
        j = i + l * d
        dn = delta(i, j)
        s, t = 0, (l + 1) >> 1
        while True:
            if delta(i, i + (s + t) * d) > dn:
                s += t
            if t == 1:
                break
            t = (t + 1) >> 1
        split = i + s * d + min(d, 0)
        lo_, hi_ = min(i, j), max(i, j)
        L = split if lo_ == split else n + split
        R = split + 1 if hi_ == split + 1 else n + split + 1
        left[n + i], right[n + i] = L, R
        parent[L] = parent[R] = n + i
    return left, right, parent, n


def area(b):
    d = np.maximum(b[3:].astype(np.float64) - b[:3].astype(np.float64), 0.0)
    return d[0] * d[1] + d[1] * d[2] + d[2] * d[0]


def ploc_tree(nb, n, radius=4):
    """PLOC (Meister & Bittner 2018) over leaves already in Morton order; nb: [2n-1, 6] boxes, leaves filled."""
    left = [-1] * (2 * n - 1); right = [-1] * (2 * n - 1); parent = [-1] * (2 * n - 1)
    C, nxt = list(range(n)), n
    while len(C) > 1:
        m = len(C)
        nn = []
        for i in range(m):
            best, arg = None, -1
            for off in range(-radius, radius + 1):
                j = i + off
                if off == 0 or j < 0 or j >= m:
                    continue
                u = np.concatenate([np.minimum(nb[C[i], :3], nb[C[j], :3]), np.maximum(nb[C[i], 3:], nb[C[j], 3:])])
                key = (area(u), min(i, j) * 0x9E3779B1 % 2**32 ^ (max(i, j) * 0x85EBCA6B % 2**32))     # (any pair-symmetric tie-break)
                if best is None or key < best:
                    best, arg = key, j
            nn.append(arg)
        out = []
        for i in range(m):
            j = nn[i]
            mutual = nn[j] == i
            if mutual and i < j:
                a, b = C[i], C[j]
                nb[nxt, :3] = np.minimum(nb[a, :3], nb[b, :3]); nb[nxt, 3:] = np.maximum(nb[a, 3:], nb[b, 3:])
                left[nxt], right[nxt] = a, b
                parent[a] = parent[b] = nxt
                out.append(nxt); nxt += 1
            elif not (mutual and i > j):
                out.append(C[i])
        assert len(out) < m, "no progress"
        C = out
    return left, right, parent, C[0]


def collapse_sizes_emit(order, nb, left, right, parent, root, n, k):
    """greedy collapse over a frontier -> kept[]; sizes bottom-up; index of every node by the ancestor formula; emission."""
    kept = [False] * (2 * n - 1)
    frontier = [root]
    while frontier:
        nxt = []
        for node in frontier:
            kept[node] = True
            kids = [left[node], right[node]]
            while len(kids) < k:
                cand = [(area(nb[c]), -p) for p, c in enumerate(kids) if c >= n]
                if not cand:
                    break
                p = -max(cand)[1]
                kids[p:p + 1] = [left[kids[p]], right[kids[p]]]
            nxt += [c for c in kids if c >= n]
        frontier = nxt
    size = [1] * (2 * n - 1)

    def fill(v):                    # iterative post-order
        stack = [(v, False)]
        while stack:
            x, done = stack.pop()
            if x < n:
                continue
            if done:
                size[x] = (1 if kept[x] else 0) + size[left[x]] + size[right[x]]
            else:
                stack += [(x, True), (left[x], False), (right[x], False)]
    fill(root)

    def index_of(v):
        idx, child, p = 0, v, parent[v]
        while p != -1:
            if kept[p]:
                idx += 1
            if right[p] == child:
                idx += size[left[p]]
            child, p = p, parent[p]
        return idx
    total = size[root]
    out = np.zeros(total, dtype=NODE_DTYPE)
    written = np.zeros(total, bool)
    for v in range(2 * n - 1):
        if v >= n and not kept[v]:
            continue
        i = index_of(v)
        assert not written[i], "two nodes at one array index"
        written[i] = True
        if v < n:
            out["w"][i, 0] = order[v]
            out["nodeData"][i] = (i + 1) | 0x80000000
        else:
            out["w"][i] = nb[v].view(np.uint32)
            out["nodeData"][i] = i + size[v]
    assert written.all()
    return out


def build(boxes, tree_type=4, quality=0):
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    n = len(boxes)
    if n == 1:
        out = np.zeros(1, dtype=NODE_DTYPE)
        out["nodeData"][0] = 1 | 0x80000000
        return out
    order, keys = morton_order(boxes)
    nb = np.zeros((2 * n - 1, 6), np.float32)
    nb[:n] = boxes[order]
    if quality == 0:
        left, right, parent, root = radix_tree(keys)
        stack = [(root, False)]        # boxes bottom-up (post-order)
        while stack:
            x, done = stack.pop()
            if x < n:
                continue
            if done:
                nb[x, :3] = np.minimum(nb[left[x], :3], nb[right[x], :3]); nb[x, 3:] = np.maximum(nb[left[x], 3:], nb[right[x], 3:])
            else:
                stack += [(x, True), (left[x], False), (right[x], False)]
    else:
        left, right, parent, root = ploc_tree(nb, n)
    return collapse_sizes_emit(order, nb, left, right, parent, root, n, tree_type)
