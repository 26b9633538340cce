"""NumPy restatement of scipy.optimize.linear_sum_assignment's rectangular LSAP solver (shortest augmenting
paths with scipy's scan order and tie rule) -- the algorithm detecting_cones_aoslo_b200/csrc/hungarian.cu runs
on the device.  Test infrastructure: pinned against the installed scipy in tests/test_host_logic.py."""
import numpy as np


def lsa_restated(cost):
    cost = np.asarray(cost, dtype=np.float64)
    nr, nc = cost.shape
    transpose = nc < nr
    if transpose:
        cost = np.ascontiguousarray(cost.T)
        nr, nc = nc, nr
    u, v = np.zeros(nr), np.zeros(nc)
    path = np.full(nc, -1, np.int64)
    col4row = np.full(nr, -1, np.int64)
    row4col = np.full(nc, -1, np.int64)
    steps = 0
    for cur in range(nr):
        remaining = np.arange(nc - 1, -1, -1)          # filled in reverse order
        num = nc
        spc = np.full(nc, np.inf)
        SR, SC = [], []
        min_val, i, sink = 0.0, cur, -1
        while sink == -1:
            steps += 1
            SR.append(i)
            js = remaining[:num]
            r = min_val + cost[i, js] - u[i] - v[js]
            upd = r < spc[js]
            path[js[upd]] = i
            spc[js[upd]] = r[upd]
            s = spc[js]
            lowest = s.min()
            cand = np.nonzero(s == lowest)[0]           # positions in scan order
            un = cand[row4col[js[cand]] == -1]
            index = un[-1] if len(un) else cand[0]      # last unassigned one, else first
            min_val = lowest
            j = js[index]
            if row4col[j] == -1:
                sink = j
            else:
                i = row4col[j]
            SC.append(j)
            num -= 1
            remaining[index] = remaining[num]
        u[cur] += min_val
        for i2 in SR:
            if i2 != cur:
                u[i2] += min_val - spc[col4row[i2]]
        for j2 in SC:
            v[j2] -= min_val - spc[j2]
        j = sink
        while True:
            i2 = path[j]
            row4col[j] = i2
            col4row[i2], j = j, col4row[i2]
            if i2 == cur:
                break
    if transpose:
        order = np.argsort(col4row, kind='stable')
        return col4row[order], order, steps
    return np.arange(nr), col4row, steps
