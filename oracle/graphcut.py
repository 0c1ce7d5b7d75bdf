"""CPU restatement of the label regularisation ``processing/generate_mesh.py:15-58`` (TEST INFRASTRUCTURE ONLY - see
``oracle/__init__.py``).

The arithmetic lives in a third-party dependency that is absent from ``/root/reference`` and from this image:
``gco-wrapper==3.0.8`` (``environment.yml:115``; Boykov-Veksler-Zabih alpha-expansion over the Boykov-Kolmogorov
max-flow).  **Parity unpinned** for this row: the published algorithm is restated - ``alpha_expansion`` performs the
expansion moves the reference's ``gc.expansion()`` performs (each move an exact minimum cut, here through
``scipy.sparse.csgraph.maximum_flow``), on the energy the reference's call sites define::

    data_cost[c, 0] = round(z[c, 1] * unary_weight),  data_cost[c, 1] = round(z[c, 0] * unary_weight)     (:25-26)
    smooth = 1 - eye(2);  set_all_neighbors(edges[:, 0], edges[:, 1], binary_weight)                          (:31-39)
    init_label_at_site(i, labels[i]);  expansion();  get_labels()                                            (:41-56)

``min_cut`` solves the same energy as one s-t cut; ``tests/test_graphcut_cpu.py`` checks that the expansion moves end
at that global minimum (two labels + Potts = submodular), which is what the device implementation computes.
"""
import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import breadth_first_order, maximum_flow


def data_costs(prediction, unary_weight):
    """int64[n,2] as the reference builds them: float32 product, round half to even, columns swapped."""
    z = np.asarray(prediction, dtype=np.float32)
    swapped = z[:, [1, 0]]
    return np.rint(swapped * np.float32(unary_weight)).astype(np.int64)


def energy(labels, cost, edges, w):
    labels = np.asarray(labels)
    data = int(cost[np.arange(cost.shape[0]), labels].sum())
    smooth = int((labels[edges[:, 0]] != labels[edges[:, 1]]).sum()) * int(w) if len(edges) else 0
    return data, smooth


def _st_cut(cap_s, cap_t, edges, w):
    """Minimum s-t cut with terminal capacities cap_s[c] (s -> c), cap_t[c] (c -> t) and symmetric arcs w[e]; returns the
    boolean mask of the cells on the SINK side."""
    n = cap_s.shape[0]
    s, t = n, n + 1
    rows = [np.full(n, s), np.arange(n), edges[:, 0], edges[:, 1]]
    cols = [np.arange(n), np.full(n, t), edges[:, 1], edges[:, 0]]
    vals = [cap_s, cap_t, w, w]
    g = sp.csr_matrix((np.concatenate(vals).astype(np.int32), (np.concatenate(rows), np.concatenate(cols))), shape=(n + 2, n + 2))
    res = maximum_flow(g, s, t)
    resid = (g - res.flow).tocsr()
    resid.data = np.where(resid.data > 0, 1, 0)
    resid.eliminate_zeros()
    reach = breadth_first_order(resid, s, directed=True, return_predecessors=False)
    on_source = np.zeros(n + 2, dtype=bool)
    on_source[reach] = True
    return ~on_source[:n], int(res.flow_value)


def min_cut(cost, edges, w):
    """Global minimiser of E(l) = sum_c cost[c, l_c] + w * #cut facets as one s-t cut (label 1 = sink side)."""
    shift = cost.min(axis=1)
    cap_s = cost[:, 1] - shift           # paid when the cell ends on the sink side (label 1)
    cap_t = cost[:, 0] - shift
    sink_side, _ = _st_cut(cap_s, cap_t, np.asarray(edges, dtype=np.int64).reshape(-1, 2), np.full(len(edges), int(w)))
    return sink_side.astype(np.int64)


def alpha_expansion(labels0, cost, edges, w, max_cycles=20):
    """gco's expansion(): cycles over alpha = 0, 1; each move is the exact minimum over "keep the label or take alpha",
    accepted if it lowers the energy; stops after a cycle without improvement."""
    labels = np.array(labels0, dtype=np.int64)
    edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    e = sum(energy(labels, cost, edges, w))
    big = int(np.abs(cost).sum() + len(edges) * w + 1)
    for _ in range(max_cycles):
        improved = False
        for alpha in (0, 1):
            # binary move variable per cell: 0 = keep, 1 = switch to alpha (cells already at alpha: both cost the same)
            keep = cost[np.arange(len(labels)), labels]
            take = cost[:, alpha]
            shift = np.minimum(keep, take)
            cap_s = take - shift                                 # paid on the sink side (= switch)
            cap_t = keep - shift
            la, lb = labels[edges[:, 0]], labels[edges[:, 1]]
            # Potts term of a move is submodular; encode exactly: pairwise cost w if the resulting labels differ
            # both at alpha already: no pairwise cost; keep-keep cost w*[la != lb] is a constant when both keep ...
            # general construction (Kolmogorov-Zabih): E(0,0)=A, E(0,1)=B, E(1,0)=C, E(1,1)=D with B + C - A - D >= 0
            A = (la != lb) * w
            B = (la != alpha) * w                                # a keeps, b switches to alpha
            C = (lb != alpha) * w
            D = np.zeros_like(A)
            cs, ct = cap_s.astype(np.int64).copy(), cap_t.astype(np.int64).copy()
            # unary parts: a: (C - A) on switching, b: (D - C) on switching; pairwise arc a -> b of B + C - A - D
            ua, ub = C - A, D - C
            np.add.at(cs, edges[:, 0], np.maximum(ua, 0)); np.add.at(ct, edges[:, 0], np.maximum(-ua, 0))
            np.add.at(cs, edges[:, 1], np.maximum(ub, 0)); np.add.at(ct, edges[:, 1], np.maximum(-ub, 0))
            pw = B + C - A - D
            n = len(labels)
            s, t = n, n + 1
            rows = np.concatenate([np.full(n, s), np.arange(n), edges[:, 0]])
            cols = np.concatenate([np.arange(n), np.full(n, t), edges[:, 1]])
            vals = np.concatenate([cs, ct, pw]).astype(np.int32)
            g = sp.csr_matrix((vals, (rows, cols)), shape=(n + 2, n + 2))
            res = maximum_flow(g, s, t)
            resid = (g - res.flow).tocsr()
            resid.data = np.where(resid.data > 0, 1, 0)
            resid.eliminate_zeros()
            reach = breadth_first_order(resid, s, directed=True, return_predecessors=False)
            src = np.zeros(n + 2, dtype=bool); src[reach] = True
            switch = ~src[:n]
            cand = np.where(switch, alpha, labels)
            ec = sum(energy(cand, cost, edges, w))
            if ec < e:
                labels, e, improved = cand, ec, True
        if not improved:
            break
    assert big > 0
    return labels
