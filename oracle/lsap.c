/* Rectangular linear-sum-assignment restated in plain C.  TEST INFRASTRUCTURE.
 *
 * PARITY PINNED AGAINST THE INSTALLED LIBRARY: the reference's association
 * step (`min_cost_matching`) calls
 * scipy.optimize.linear_sum_assignment, whose source is not in /root/reference
 * (third-party, scipy 1.18.1 installed here).  This file restates the
 * published algorithm (D. F. Crouse, "On implementing 2D rectangular
 * assignment algorithms", IEEE TAES 2016) including the
 * library's observable tie-break rules:
 *   - tall matrices (nr > nc) are solved transposed, result re-sorted by row;
 *   - `remaining[]` is filled in reverse column order;
 *   - among equal shortest-path costs an unassigned column wins, otherwise the
 *     first-scanned one is kept (strict `<`);
 *   - all arithmetic in double, expression order
 *     ((minVal + cost[i][j]) - u[i]) - v[j].
 * tests/test_oracle_lsap.py checks it against scipy on random, integer-tied
 * and constant matrices.  The CUDA kernel (csrc/lsap.cu) mirrors this file.
 *
 * Returns 0 on success, -1 if infeasible.  row4match[k], col4match[k] for
 * k < min(nr,nc), rows ascending.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int ssb_oracle_lsap(const double *cost_in, int nr, int nc, int64_t *rows_out,
                    int64_t *cols_out)
{
    if (nr == 0 || nc == 0) return 0;
    int transpose = nc < nr;
    double *cost = (double *)malloc(sizeof(double) * (size_t)nr * nc);
    if (transpose) {
        for (int i = 0; i < nr; i++)
            for (int j = 0; j < nc; j++) cost[(size_t)j * nr + i] = cost_in[(size_t)i * nc + j];
        int t = nr; nr = nc; nc = t;
    } else {
        memcpy(cost, cost_in, sizeof(double) * (size_t)nr * nc);
    }
    double *u = (double *)calloc(nr, sizeof(double));
    double *v = (double *)calloc(nc, sizeof(double));
    double *spc = (double *)malloc(sizeof(double) * nc);
    int *path = (int *)malloc(sizeof(int) * nc);
    int *col4row = (int *)malloc(sizeof(int) * nr);
    int *row4col = (int *)malloc(sizeof(int) * nc);
    char *SR = (char *)malloc(nr), *SC = (char *)malloc(nc);
    int *remaining = (int *)malloc(sizeof(int) * nc);
    for (int i = 0; i < nr; i++) col4row[i] = -1;
    for (int j = 0; j < nc; j++) { row4col[j] = -1; path[j] = -1; }
    int rc = 0;

    for (int curRow = 0; curRow < nr && rc == 0; curRow++) {
        double minVal = 0;
        int i = curRow;
        int num_remaining = nc;
        for (int it = 0; it < nc; it++) remaining[it] = nc - it - 1;
        memset(SR, 0, nr); memset(SC, 0, nc);
        for (int j = 0; j < nc; j++) spc[j] = INFINITY;
        int sink = -1;
        while (sink == -1) {
            int index = -1;
            double lowest = INFINITY;
            SR[i] = 1;
            for (int it = 0; it < num_remaining; it++) {
                int j = remaining[it];
                double r = minVal + cost[(size_t)i * nc + j] - u[i] - v[j];
                if (r < spc[j]) { path[j] = i; spc[j] = r; }
                if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) {
                    lowest = spc[j];
                    index = it;
                }
            }
            minVal = lowest;
            if (minVal == INFINITY) { rc = -1; break; }
            int j = remaining[index];
            if (row4col[j] == -1) sink = j; else i = row4col[j];
            SC[j] = 1;
            remaining[index] = remaining[--num_remaining];
        }
        if (rc) break;
        u[curRow] += minVal;
        for (int r = 0; r < nr; r++)
            if (SR[r] && r != curRow) u[r] += minVal - spc[col4row[r]];
        for (int j = 0; j < nc; j++)
            if (SC[j]) v[j] -= minVal - spc[j];
        int j = sink;
        while (1) {
            int r = path[j];
            row4col[j] = r;
            int t = col4row[r]; col4row[r] = j; j = t;
            if (r == curRow) break;
        }
    }
    if (rc == 0) {
        if (transpose) {
            /* col4row is indexed by original column; emit sorted by original row
             * (values of col4row are distinct, so a counting pass suffices) */
            int k = 0;
            for (int r = 0; r < nc; r++) {        /* nc == original nr */
                int c = row4col[r];                /* original col matched to original row r */
                if (c >= 0) { rows_out[k] = r; cols_out[k] = c; k++; }
            }
        } else {
            for (int r = 0; r < nr; r++) { rows_out[r] = r; cols_out[r] = col4row[r]; }
        }
    }
    free(cost); free(u); free(v); free(spc); free(path); free(col4row);
    free(row4col); free(SR); free(SC); free(remaining);
    return rc;
}
