"""Generates tests/golden/ising_4x4_exact.json: exact enumeration of the 2^16 states of the 4x4 Ising model
(J = 1, H = 0), physical energy convention (every bond once).  Pure numpy; no reference code involved.
Observables as the reference's StatSensor defines them (src/instrument.rs:110-131, src/accumulator.rs:50-63):
Cv = Var(E)/(N T^2), chi = Var(|M|)/(N T), U4 = 1 - <M^4>/(3 <M^2>^2), E and |M| totals."""
import json
import numpy as np


def table(pbc):
    L = 4
    idx = np.arange(16).reshape(4, 4)
    bonds = []
    for y in range(L):
        for x in range(L):
            if x + 1 < L or pbc: bonds.append((idx[y, x], idx[y, (x + 1) % L]))
            if y + 1 < L or pbc: bonds.append((idx[y, x], idx[(y + 1) % L, x]))
    s = ((np.arange(1 << 16)[:, None] >> np.arange(16)) & 1) * 2 - 1
    E = np.zeros(1 << 16)
    for a, b in bonds:
        E -= s[:, a] * s[:, b]
    M = np.abs(s.sum(axis=1)).astype(float)
    rows = []
    for T in (1.5, 2.0, 2.269185314213022, 3.0, 4.0):
        w = np.exp(-(E - E.min()) / T); w /= w.sum()
        e1, e2 = (w * E).sum(), (w * E * E).sum()
        m1, m2, m4 = (w * M).sum(), (w * M * M).sum(), (w * M ** 4).sum()
        rows.append(dict(T=T, E=e1, Cv=(e2 - e1 * e1) / (16 * T * T), M=m1, chi=(m2 - m1 * m1) / (16 * T), U4=1 - m4 / (3 * m2 * m2)))
    return dict(n_bonds=len(bonds), rows=rows)


if __name__ == "__main__":
    out = dict(pbc=table(True), open=table(False))
    with open(__file__.replace("make_exact_enumeration.py", "ising_4x4_exact.json"), "w") as f:
        json.dump(out, f, indent=1)
    for k, v in out.items():
        print(k, v["n_bonds"], [round(r["E"], 10) for r in v["rows"]])
