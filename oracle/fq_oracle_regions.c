/* Test oracle (never linked into the product): RegionList of the reference, restated for ONE chromosome over plain
 * arrays -- AddRegion (src/RegionList.cpp:65-74: map[start] = end, a later region with the same start replaces the
 * earlier one), Collapse (76-118), Join(b, isUnion = false) (120-173) and IsOverlapped (47-63).
 * A region list is an array of (start, end) pairs kept sorted by start, as std::map<int,int> iterates. */
#include <stdlib.h>
#include <string.h>

#include "fq_oracle.h"

static int reg_add(int *r, int n, int start, int end)          /* map[start] = end */
{
    int i, j;
    for (i = 0; i < n; ++i) {
        if (r[2 * i] == start) { r[2 * i + 1] = end; return n; }
        if (r[2 * i] > start) break;
    }
    for (j = n; j > i; --j) { r[2 * j] = r[2 * j - 2]; r[2 * j + 1] = r[2 * j - 1]; }
    r[2 * i] = start; r[2 * i + 1] = end;
    return n + 1;
}

/* Collapse: returns the new count, *len = total length */
int orc_regions_collapse(int *r, int n, long long *len)
{
    int *t = (int *)malloc((size_t)(2 * n + 2) * sizeof(int)), nt = 0, holder = 0, i;
    long long l = 0;
    if (n == 0) { free(t); if (len) *len = 0; return 0; }
    for (i = 0; i < n; ++i) {
        int beg1 = r[2 * holder], end1 = r[2 * holder + 1], beg2 = r[2 * i], end2 = r[2 * i + 1];
        if (end1 >= end2) continue;
        else if (end1 < beg2) { nt = reg_add(t, nt, beg1, end1); holder = i; }
        else { nt = reg_add(t, nt, beg1, end2); r[2 * holder + 1] = end2; }
    }
    nt = reg_add(t, nt, r[2 * holder], r[2 * holder + 1]);
    memcpy(r, t, (size_t)(2 * nt) * sizeof(int));
    free(t);
    for (i = 0; i < nt; ++i) l += (r[2 * i + 1] - r[2 * i]) + 1;
    if (len) *len = l;
    return nt;
}

/* a.Join(b, false): a is collapsed, intersected with b (as given), collapsed again; a is overwritten */
int orc_regions_join(int *a, int na, const int *b, int nb, long long *len)
{
    int *t = (int *)malloc((size_t)(2 * (na + nb) + 2) * sizeof(int)), nt = 0, i = 0, j = 0;
    na = orc_regions_collapse(a, na, 0);
    while (i < na && j < nb) {
        int beg1 = a[2 * i], end1 = a[2 * i + 1], beg2 = b[2 * j], end2 = b[2 * j + 1];
        if (beg1 <= beg2) {
            if (end1 > end2) { nt = reg_add(t, nt, beg2, end2); ++j; }
            else if (end1 > beg2) { nt = reg_add(t, nt, beg2, end1); ++i; }
            else ++i;
        } else {
            if (end1 <= end2) { nt = reg_add(t, nt, beg1, end1); ++i; }
            else if (end1 > beg2 && beg1 < end2) { nt = reg_add(t, nt, beg1, end2); ++j; }
            else ++j;
        }
    }
    memcpy(a, t, (size_t)(2 * nt) * sizeof(int));
    free(t);
    return orc_regions_collapse(a, nt, len);
}

int orc_regions_add(int *r, int n, int start, int end) { return reg_add(r, n, start, end); }

int orc_regions_overlapped(const int *r, int n, int pos)
{
    int i = 0;
    while (i < n && r[2 * i] < pos) ++i;              /* lower_bound(pos) */
    if (i < n && r[2 * i] <= pos && r[2 * i + 1] >= pos) return 1;
    if (i > 0) { --i; if (r[2 * i] <= pos && r[2 * i + 1] >= pos) return 1; }
    return 0;
}
