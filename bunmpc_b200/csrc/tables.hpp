// Symbolic tables for the two FISTA sub-problems (host side, built once per (n_col, n_eff)).
//
// The constraint matrices of the reference have a FIXED sparsity pattern:
//   A_x (force problem)  centroidal.cpp:67-82   -- 9 e n structural entries
//   A_f (state problem)  centroidal.cpp:14-25,89-100, centroidal.hpp:22-27 -- 27 n + 9 structural entries
// so everything Eigen derives from the pattern at run time (A^T A, A^T b, row/column traversal order,
// problem.cpp:31-39,46-56) is derived here once, as index tables that the kernel walks in the
// reference's accumulation order (ascending row k for A^T A / A^T b, ascending column for A y).
//
// Entry values live in a compact per-instance array in shared memory ("aidx" space):
//   A_x:  aidx = 9 (e t + f) + q ; q = 0..2 velocity rows, 3..8 = (6,by)(6,bz)(7,bx)(7,bz)(8,bx)(8,by)
//   A_f:  aidx = 27 t + {0..8 diag(+1), 9..17 next(-1), 18..20 dt, 21..26 cross}, 27 n + k = x_init rows
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace bunmpc {

struct Entry { int row, col, aidx; };

struct HostTables {
    int nv = 0, nr = 0, nvp = 0, nrp = 0, nval = 0;
    int KH = 0, PM = 0, KA = 0, KC = 0;      // maxima found (must not exceed the kernel's template bounds)
    bool contiguous_rows = true;             // every row of A^T A occupies consecutive columns
    std::vector<uint8_t> h_len, h_np, c_len, a_len;
    std::vector<uint16_t> h_col, c_row, c_aidx, a_col, a_aidx;
    std::vector<uint32_t> h_pair;
};

inline std::vector<Entry> pattern_Ax(int n, int e)
{
    std::vector<Entry> E;
    static const int cr_r[6] = {6, 6, 7, 7, 8, 8}, cr_c[6] = {1, 2, 0, 2, 0, 1};
    for (int t = 0; t < n; ++t)
        for (int f = 0; f < e; ++f) {
            int base = 3 * e * t + 3 * f, a0 = 9 * (e * t + f);
            for (int k = 0; k < 3; ++k) E.push_back({9 * t + 3 + k, base + k, a0 + k});
            for (int q = 0; q < 6; ++q) E.push_back({9 * t + cr_r[q], base + cr_c[q], a0 + 3 + q});
        }
    return E;
}

inline std::vector<Entry> pattern_Af(int n)
{
    std::vector<Entry> E;
    static const int cr_r[6] = {6, 6, 7, 7, 8, 8}, cr_c[6] = {1, 2, 0, 2, 0, 1};
    for (int t = 0; t < n; ++t) {
        for (int l = 0; l < 9; ++l) {
            E.push_back({9 * t + l, 9 * t + l, 27 * t + l});
            E.push_back({9 * t + l, 9 * (t + 1) + l, 27 * t + 9 + l});
        }
        for (int l = 0; l < 3; ++l) E.push_back({9 * t + l, 9 * (t + 1) + l + 3, 27 * t + 18 + l});
        for (int q = 0; q < 6; ++q) E.push_back({9 * t + cr_r[q], 9 * t + cr_c[q], 27 * t + 21 + q});
    }
    for (int k = 0; k < 9; ++k) E.push_back({9 * n + k, k, 27 * n + k});
    return E;
}

// Tables in [slot][index] layout (coalesced when thread <-> index), leading dimension padded to 32.
// Unused slots are PADDED so that the kernel can walk every table without branches: value indices point at
// `zero_aidx` (an always-zero element of the value array), columns at `zero_col` (an always-zero element of the
// iterate vectors), rows at row 0 (multiplied by a zero value).
inline HostTables build_tables(const std::vector<Entry> &E, int nr, int nv, int nval,
                               int KH, int PM, int KA, int KC, int zero_aidx, int zero_col)
{
    HostTables T;
    T.nv = nv; T.nr = nr; T.nval = nval;
    T.nvp = (nv + 31) / 32 * 32; T.nrp = (nr + 31) / 32 * 32;
    std::vector<std::vector<Entry>> byCol(nv), byRow(nr);
    for (const Entry &x : E) { byCol[x.col].push_back(x); byRow[x.row].push_back(x); }
    for (auto &v : byCol) std::sort(v.begin(), v.end(), [](const Entry &a, const Entry &b) { return a.row < b.row; });
    for (auto &v : byRow) std::sort(v.begin(), v.end(), [](const Entry &a, const Entry &b) { return a.col < b.col; });

    T.h_len.assign(T.nvp, 0); T.c_len.assign(T.nvp, 0); T.a_len.assign(T.nrp, 0);
    const uint32_t zpair = (uint32_t)zero_aidx | ((uint32_t)zero_aidx << 16);
    T.h_col.assign((size_t)KH * T.nvp, (uint16_t)zero_col); T.h_np.assign((size_t)KH * T.nvp, 0);
    T.h_pair.assign((size_t)KH * PM * T.nvp, zpair);
    T.c_row.assign((size_t)KC * T.nvp, 0); T.c_aidx.assign((size_t)KC * T.nvp, (uint16_t)zero_aidx);
    T.a_col.assign((size_t)KA * T.nrp, (uint16_t)zero_col); T.a_aidx.assign((size_t)KA * T.nrp, (uint16_t)zero_aidx);

    for (int r = 0; r < nr; ++r) {                       // rows of A, ascending column
        int len = (int)byRow[r].size();
        T.KA = std::max(T.KA, len);
        if (len > KA) continue;
        T.a_len[r] = (uint8_t)len;
        for (int q = 0; q < len; ++q) {
            T.a_col[(size_t)q * T.nrp + r] = (uint16_t)byRow[r][q].col;
            T.a_aidx[(size_t)q * T.nrp + r] = (uint16_t)byRow[r][q].aidx;
        }
    }
    std::vector<char> mark(nv, 0);
    for (int i = 0; i < nv; ++i) {
        const auto &ci = byCol[i];                       // column i of A, ascending row
        int clen = (int)ci.size();
        T.KC = std::max(T.KC, clen);
        if (clen <= KC) {
            T.c_len[i] = (uint8_t)clen;
            for (int p = 0; p < clen; ++p) {
                T.c_row[(size_t)p * T.nvp + i] = (uint16_t)ci[p].row;
                T.c_aidx[(size_t)p * T.nvp + i] = (uint16_t)ci[p].aidx;
            }
        }
        // row i of A^T A: every column j sharing a row with column i (plus the diagonal), ascending j
        mark[i] = 1;
        for (const Entry &x : ci) for (const Entry &y : byRow[x.row]) mark[y.col] = 1;
        int k = 0;
        for (int j = 0; j < nv; ++j) {
            if (!mark[j]) continue;
            mark[j] = 0;
            const auto &cj = byCol[j];
            std::vector<uint32_t> pairs;                 // shared rows, ascending
            size_t a = 0, b = 0;
            while (a < ci.size() && b < cj.size()) {
                if (ci[a].row < cj[b].row) ++a;
                else if (ci[a].row > cj[b].row) ++b;
                else { pairs.push_back((uint32_t)ci[a].aidx | ((uint32_t)cj[b].aidx << 16)); ++a; ++b; }
            }
            T.PM = std::max(T.PM, (int)pairs.size());
            if (k < KH && (int)pairs.size() <= PM) {
                T.h_col[(size_t)k * T.nvp + i] = (uint16_t)j;
                T.h_np[(size_t)k * T.nvp + i] = (uint8_t)pairs.size();
                for (size_t p = 0; p < pairs.size(); ++p)
                    T.h_pair[((size_t)k * PM + p) * T.nvp + i] = pairs[p];
            }
            ++k;
        }
        T.KH = std::max(T.KH, k);
        if (k <= KH) T.h_len[i] = (uint8_t)k;
        for (int q = 1; q < k && q < KH; ++q)
            if (T.h_col[(size_t)q * T.nvp + i] != T.h_col[i] + q) T.contiguous_rows = false;
    }
    return T;
}

}  // namespace bunmpc
