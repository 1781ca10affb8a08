// lattice.hpp -- host-side lattice description for the product: unit-cell tables, neighbour entries
// for the implicit device stencil, colouring rules and the reference-form CSR export.
//
// Restates what `Lattice::{sc,bcc,fcc}(1.0).expand(x,y,z).drop_*()` (src/input.rs:296-322) and
// Exchange::from_lattice (src/energy.rs:176-187) produce.  vegas-lattice 0.13 is not available,
// so the unit-cell conventions are this project's own and are documented in DESIGN.md
// ("parity unpinned").  Independent of oracle/ (the tests compare the two).
#pragma once
#include <cstdint>
#include <vector>
#include <algorithm>

namespace vgl {

struct UcEdge { int s, t, dx, dy, dz; };

inline const std::vector<UcEdge>& unitcell_edges(int unitcell) {
    static const std::vector<UcEdge> sc = {{0, 0, 1, 0, 0}, {0, 0, 0, 1, 0}, {0, 0, 0, 0, 1}};
    static std::vector<UcEdge> bcc, fcc;
    if (bcc.empty()) {
        for (int dz = 0; dz >= -1; --dz)
            for (int dy = 0; dy >= -1; --dy)
                for (int dx = 0; dx >= -1; --dx) bcc.push_back({0, 1, dx, dy, dz});
        // A(0,0,0) B(1/2,1/2,0) C(1/2,0,1/2) D(0,1/2,1/2): the two half-integer axes of each pair carry the offsets
        auto pair = [&](int s, int t, int ax0, int lo0, int hi0, int ax1, int lo1, int hi1) {
            for (int b = lo1; b <= hi1; ++b)
                for (int a = lo0; a <= hi0; ++a) {
                    int d[3] = {0, 0, 0};
                    d[ax0] = a; d[ax1] = b;
                    fcc.push_back({s, t, d[0], d[1], d[2]});
                }
        };
        pair(0, 1, 0, -1, 0, 1, -1, 0);  // A-B: dx,dy in {0,-1}
        pair(0, 2, 0, -1, 0, 2, -1, 0);  // A-C: dx,dz
        pair(0, 3, 1, -1, 0, 2, -1, 0);  // A-D: dy,dz
        pair(1, 2, 1, 0, 1, 2, -1, 0);   // B-C: dy in {0,1}, dz in {0,-1}
        pair(1, 3, 0, 0, 1, 2, -1, 0);   // B-D: dx in {0,1}, dz in {0,-1}
        pair(2, 3, 0, 0, 1, 1, -1, 0);   // C-D: dx in {0,1}, dy in {0,-1}
    }
    return unitcell == 0 ? sc : (unitcell == 1 ? bcc : fcc);
}

inline int basis_count(int unitcell) { return unitcell == 0 ? 1 : (unitcell == 1 ? 2 : 4); }

struct Desc {
    int unitcell;
    uint64_t nx, ny, nz;
    int pbc[3];
    int literal;
};

// Calls f(i, j) for every entry of row i of the reference CSR *before* duplicate merging.
template <typename F>
inline void for_each_neighbour(const Desc& d, uint64_t i, F&& f) {
    const int nb = basis_count(d.unitcell);
    const uint64_t cell = i / nb;
    const int b = (int)(i % nb);
    const int64_t ix = (int64_t)(cell % d.nx), iy = (int64_t)((cell / d.nx) % d.ny), iz = (int64_t)(cell / (d.nx * d.ny));
    const int64_t L[3] = {(int64_t)d.nx, (int64_t)d.ny, (int64_t)d.nz};
    for (const UcEdge& e : unitcell_edges(d.unitcell)) {
        for (int dir = 0; dir < 2; ++dir) {  // dir 0: i is the source; 1: i is the target
            if (dir == 0 ? e.s != b : e.t != b) continue;
            int64_t t[3] = {ix + (dir ? -e.dx : e.dx), iy + (dir ? -e.dy : e.dy), iz + (dir ? -e.dz : e.dz)};
            bool drop = false;
            for (int a = 0; a < 3; ++a)
                if (t[a] < 0 || t[a] >= L[a]) {
                    if (!d.pbc[a]) { drop = true; break; }
                    t[a] = (t[a] % L[a] + L[a]) % L[a];
                }
            if (drop) continue;
            const uint64_t j = (((uint64_t)t[2] * d.ny + (uint64_t)t[1]) * d.nx + (uint64_t)t[0]) * nb + (dir ? e.s : e.t);
            if (d.literal && (dir == 0 ? !(i <= j) : !(j <= i))) continue;
            f(i, j);
        }
    }
}

// Reference-form CSR: sorted columns, duplicates summed (sprs TriMat::to_csr).
inline void build_csr(const Desc& d, double J, std::vector<uint64_t>& row_ptr, std::vector<uint32_t>& col,
                      std::vector<double>& val) {
    const uint64_t n = d.nx * d.ny * d.nz * basis_count(d.unitcell);
    row_ptr.assign(n + 1, 0);
    col.clear(); val.clear();
    std::vector<uint32_t> tmp;
    for (uint64_t i = 0; i < n; ++i) {
        tmp.clear();
        for_each_neighbour(d, i, [&](uint64_t, uint64_t j) { tmp.push_back((uint32_t)j); });
        std::sort(tmp.begin(), tmp.end());
        for (size_t k = 0; k < tmp.size(); ++k) {
            if (k > 0 && tmp[k] == tmp[k - 1]) val.back() += J;
            else { col.push_back(tmp[k]); val.push_back(J); }
        }
        row_ptr[i + 1] = col.size();
    }
}

// Colour of a site of a structured lattice; n_colours set by colour_count().
// bcc/fcc: basis index.  sc: parity when every periodic axis that still carries wrap bonds is even,
// else a mod-3 scheme (the last coordinate of an odd periodic axis gets the third value).
struct Colouring {
    int n_colours;
    int mod3;
    Desc d;
    explicit Colouring(const Desc& dd) : d(dd) {
        if (d.unitcell != 0) { n_colours = basis_count(d.unitcell); mod3 = 0; return; }
        const uint64_t L[3] = {d.nx, d.ny, d.nz};
        mod3 = 0;
        for (int a = 0; a < 3; ++a)
            if (d.pbc[a] && !d.literal && L[a] >= 3 && (L[a] & 1)) mod3 = 1;
        n_colours = mod3 ? 3 : 2;
        if (d.nx * d.ny * d.nz == 1) n_colours = 1;
    }
    int colour(uint64_t i) const {
        if (d.unitcell != 0) return (int)(i % basis_count(d.unitcell));
        if (n_colours == 1) return 0;
        const uint64_t x = i % d.nx, y = (i / d.nx) % d.ny, z = i / (d.nx * d.ny);
        if (!mod3) return (int)((x + y + z) & 1);
        const uint64_t L[3] = {d.nx, d.ny, d.nz}, c[3] = {x, y, z};
        uint64_t s = 0;
        for (int a = 0; a < 3; ++a) {
            const bool odd_wrap = d.pbc[a] && L[a] >= 3 && (L[a] & 1);
            s += (odd_wrap && c[a] == L[a] - 1) ? 2 : (c[a] & 1);
        }
        return (int)(s % 3);
    }
};

// first-fit greedy colouring of a CSR graph (diagonal ignored); returns number of colours, -1 if > 64.
// Exchange::new accepts any CsMat (src/energy.rs:171-173), also one whose pattern is not symmetric: a colour class must
// be an independent set of the SYMMETRISED pattern (if j lists i but i does not list j, a pass that updates i while j
// reads it would race), so the transpose's entries count as neighbours too.
inline int greedy_colour(uint64_t n, const uint64_t* row_ptr, const uint32_t* col, std::vector<uint8_t>& colour) {
    std::vector<uint64_t> tptr(n + 1, 0);
    for (uint64_t p = 0; p < row_ptr[n]; ++p) tptr[col[p] + 1]++;
    for (uint64_t i = 0; i < n; ++i) tptr[i + 1] += tptr[i];
    std::vector<uint32_t> tcol(row_ptr[n]);
    {
        std::vector<uint64_t> fill(tptr.begin(), tptr.end() - 1);
        for (uint64_t i = 0; i < n; ++i)
            for (uint64_t p = row_ptr[i]; p < row_ptr[i + 1]; ++p) tcol[fill[col[p]]++] = (uint32_t)i;
    }
    colour.assign(n, 0xFF);
    int ncol = 0;
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t used = 0;
        for (uint64_t p = row_ptr[i]; p < row_ptr[i + 1]; ++p) {
            const uint32_t j = col[p];
            if (j != i && colour[j] != 0xFF) used |= 1ull << colour[j];
        }
        for (uint64_t p = tptr[i]; p < tptr[i + 1]; ++p) {
            const uint32_t j = tcol[p];
            if (j != i && colour[j] != 0xFF) used |= 1ull << colour[j];
        }
        int c = 0;
        while (c < 64 && (used >> c & 1)) ++c;
        if (c >= 64) return -1;
        colour[i] = (uint8_t)c;
        ncol = std::max(ncol, c + 1);
    }
    return ncol;
}

}  // namespace vgl
