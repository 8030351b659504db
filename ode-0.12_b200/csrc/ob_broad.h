// ob_broad.h — order-exact emulation of dxHashSpace::collide
// (ode/src/collision_space.cpp:420-583) without hash tables or pointer chasing.
//
// The reference walks every AABB through a chained hash table and reports each
// overlapping pair the FIRST time it is encountered.  That first encounter is
// a closed-form function of the pair:
//   * geoms are numbered by their position in the space list ("walk index" w,
//     head = 0); the reference's first_aabb list is the reverse walk order, so
//     the query rank of a hashed geom is  (n_hashed-1 - hashed_rank(w)).
//   * query q finds node p only at level L = level(p) >= level(q), in the
//     first cell (x,y,z ascending scan) that both cover at that level, i.e.
//     the component-wise max of the two lower cell corners, and within one
//     cell nodes are met in walk order.
//   * if both levels are equal the pair is met first by whichever of the two
//     is queried first (the one with the LARGER walk index).
// So every candidate pair gets the lexicographic key
//   (stage, query rank, level, cx, cy, cz, node walk index)
// and sorting the surviving pairs by key reproduces the callback order exactly.
// Stage 1 = hashed x big-box list (:569-573), stage 2 = big x big (:576-580).
// Hash collisions between different cells cannot reorder same-cell nodes, so
// the table size / getVirtualAddress never enter the result.
#pragma once
#include "ob_types.h"

#define OB_LEVEL_BIG 0x7fffffff

struct ObCellBox {  // per geom, per step
  int level;        // OB_LEVEL_BIG -> big_boxes list
  int db[6];        // discretized bounds at `level`
};

// findLevel (:331-351) + level clamp + discretization (:454-462)
OB_HD void ob_hash_cellbox(const real *aabb, int minlevel, int maxlevel, ObCellBox *out) {
  if (aabb[0] <= -OB_INF || aabb[1] >= OB_INF || aabb[2] <= -OB_INF || aabb[3] >= OB_INF || aabb[4] <= -OB_INF ||
      aabb[5] >= OB_INF) {
    out->level = OB_LEVEL_BIG;
    return;
  }
  real q = aabb[1] - aabb[0], q2 = aabb[3] - aabb[2];
  if (q2 > q) q = q2;
  q2 = aabb[5] - aabb[4];
  if (q2 > q) q = q2;
  int level;
  // the reference calls the double frexp on a dReal (collision_space.cpp:349); glibc
  // stores exponent 0 for NaN/Inf arguments, which CUDA's frexp does not promise
  if (q - q == 0) frexp((double)q, &level); else level = 0;
  if (level < minlevel) level = minlevel;
  if (level > maxlevel) { out->level = OB_LEVEL_BIG; return; }
  out->level = level;
  real cellsize = (real)ldexp(1.0, level);
  for (int i = 0; i < 6; i++) {
    // (int) of a NaN / out-of-range double is INT_MIN on x86-64 (cvttsd2si); make that explicit
    double f = floor((double)(aabb[i] / cellsize));
    out->db[i] = (f >= -2147483648.0 && f < 2147483648.0) ? (int)f : (int)0x80000000;
  }
}

struct ObPairKey {
  int k[7];   // stage, query rank, level, cx, cy, cz, node walk index
};

OB_HD bool ob_key_less(const ObPairKey &a, const ObPairKey &b) {
  for (int i = 0; i < 7; i++) {
    if (a.k[i] < b.k[i]) return true;
    if (a.k[i] > b.k[i]) return false;
  }
  return false;
}

// collideAABBs filter (ode/src/collision_space_internal.h:48-82) minus the callback
OB_HD bool ob_aabb_pair_filter(int body1, int body2, uint32_t cat1, uint32_t col1, uint32_t cat2, uint32_t col2,
                               const real *b1, const real *b2) {
  if (body1 == body2 && body1 >= 0) return false;
  if (((cat1 & col2) || (cat2 & col1)) == 0) return false;
  if (b1[0] > b2[1] || b1[1] < b2[0] || b1[2] > b2[3] || b1[3] < b2[2] || b1[4] > b2[5] || b1[5] < b2[4]) return false;
  return true;
}

// Given two geoms by walk index (wa < wb), their cell boxes and their ranks
// (hashed rank hr = number of hashed geoms with smaller walk index; big rank
// likewise among big geoms), produce the key and the (o1,o2) orientation.
// nh = number of hashed geoms, nbig = number of big geoms.
// Returns false when the two cell boxes never meet in the table (only possible
// for NaN / overflowed AABBs, whose pairs pass the float AABB test vacuously).
OB_HD bool ob_hash_pair_key(int wa, int wb, const ObCellBox &ca, const ObCellBox &cb, int hra, int hrb, int bra,
                            int brb, int nh, int nbig, ObPairKey *key, int *first_is_a) {
  bool biga = ca.level == OB_LEVEL_BIG, bigb = cb.level == OB_LEVEL_BIG;
  if (!biga && !bigb) {
    // query = lower level; on a tie the one queried first = larger walk index = b
    bool q_is_a = ca.level < cb.level;
    const ObCellBox &Q = q_is_a ? ca : cb;
    const ObCellBox &P = q_is_a ? cb : ca;
    int sh = P.level - Q.level;
    if (sh > 31) sh = 31;
    int qx = Q.db[0] >> sh, qy = Q.db[2] >> sh, qz = Q.db[4] >> sh;   // arithmetic shift == repeated >>=1 (:562)
    int qx1 = Q.db[1] >> sh, qy1 = Q.db[3] >> sh, qz1 = Q.db[5] >> sh;
    if (qx > P.db[1] || qx1 < P.db[0] || qy > P.db[3] || qy1 < P.db[2] || qz > P.db[5] || qz1 < P.db[4]) return false;
    key->k[0] = 0;
    key->k[1] = nh - 1 - (q_is_a ? hra : hrb);
    key->k[2] = P.level;
    key->k[3] = qx > P.db[0] ? qx : P.db[0];
    key->k[4] = qy > P.db[2] ? qy : P.db[2];
    key->k[5] = qz > P.db[4] ? qz : P.db[4];
    key->k[6] = q_is_a ? wb : wa;
    *first_is_a = q_is_a;
  } else if (biga != bigb) {
    // normal x big: outer = first_aabb order, inner = big_boxes order (reverse walk)
    bool n_is_a = bigb;
    key->k[0] = 1;
    key->k[1] = nh - 1 - (n_is_a ? hra : hrb);
    key->k[2] = nbig - 1 - (n_is_a ? brb : bra);
    key->k[3] = key->k[4] = key->k[5] = key->k[6] = 0;
    *first_is_a = n_is_a;
  } else {
    // big x big: big_boxes list is reverse walk order, pairs (i, later j) -> first = larger walk index
    key->k[0] = 2;
    key->k[1] = nbig - 1 - brb;
    key->k[2] = nbig - 1 - bra;
    key->k[3] = key->k[4] = key->k[5] = key->k[6] = 0;
    *first_is_a = 0;
  }
  return true;
}

// ---- dxSAPSpace::collide (collision_sapspace.cpp:425-496) and dxSimpleSpace::collide -------------
// (collision_space.cpp:247-268), order-exact.
//
// SAP: TmpGeomList = enabled geoms with a finite axis-0 maximum, in GeomList order; the others go to
// TmpInfGeomList.  BoxPruning (:521-567) sorts float-cast axis-0 minima (+ an FLT_MAX sentinel) with
// RadixSort (:600-830) and sweeps.  RadixSort's output, as a function of its inputs:
//   * it starts from the permutation the previous call returned (or identity when that is invalid:
//     first call, or the element count changed, :109-129);
//   * if the keys read in that order are already non-decreasing it returns that order unchanged (:633-687);
//   * otherwise the result is the stable LSB-radix order: by IEEE bit pattern, negatives (descending
//     bits) before non-negatives (ascending bits), equal keys in starting order EXCEPT equal negative
//     keys, which come out reversed (the MSB pass writes negatives back to front, :751-824).
// The sweep visits sorted position k, then every later position j while key[j] <= max0[k]
// (:537-566), so pair (k, j) has sequence key (0, k, j).  Infinite geoms: (1, m, 0, n) against later
// infinite ones, then (1, m, 1, t) against every finite one in TmpGeomList order (:478-493).
OB_HD uint32_t ob_sap_keyorder(float f) {
  const uint32_t u = (uint32_t)ob_f2i(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
// does element u come before element t in RadixSort's output (full-sort case)?
OB_HD bool ob_sap_precedes(uint32_t ou, uint32_t ot, int init_u, int init_t) {
  if (ou != ot) return ou < ot;
  return (ou & 0x80000000u) ? init_u < init_t : init_u > init_t;   // equal negative keys: reversed
}
OB_HD void ob_sap_axes(int axisorder, int *ax0, int *ax1, int *ax2) {
  *ax0 = ((axisorder) & 3) << 1; *ax1 = ((axisorder >> 2) & 3) << 1; *ax2 = ((axisorder >> 4) & 3) << 1;
}
// collideGeomsNoAABBs filter (:234-258) minus the callback
OB_HD bool ob_pair_filter_noaabb(int body1, int body2, uint32_t cat1, uint32_t col1, uint32_t cat2, uint32_t col2) {
  if (body1 == body2 && body1 >= 0) return false;
  return ((cat1 & col2) || (cat2 & col1)) != 0;
}
// finite x finite: K = geom at the earlier sorted position, J = the later one
OB_HD bool ob_sap_sweep_test(float key_j, const real *aabb_k, const real *aabb_j, int ax0, int ax1, int ax2) {
  if (!((real)key_j <= aabb_k[ax0 + 1])) return false;
  if (!(aabb_k[ax1 + 1] >= aabb_j[ax1] && aabb_j[ax1 + 1] >= aabb_k[ax1])) return false;
  if (!(aabb_k[ax2 + 1] >= aabb_j[ax2] && aabb_j[ax2 + 1] >= aabb_k[ax2])) return false;
  return true;
}
