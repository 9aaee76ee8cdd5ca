// vce_orient.h -- K1: genotype / phase expansion of one variant into its ordered list of orientations.
//
// Restates Orientor.orientations for every orientor of the reference
// (src/com/rtg/vcf/eval/Orientor.java:60-338, util/Permutation.java:43-107).  The ORDER of the rows is
// significant: PathFinder inserts child paths in this order (Path.java:311-320).
// One thread handles one variant; the same code counts rows (first pass) and writes them (second pass).
#ifndef VCE_ORIENT_H_
#define VCE_ORIENT_H_

#include "vce_device.h"
#include "../../include/vce.h"

namespace vce {

struct VariantGt {
  int gtPloidy;       // entries valid in gt[]
  int gt[VCE_MAX_PLOIDY];
  int phased;
  int slot0;          // first allele slot of the variant (slot0 + id + 1 = allele id)
  int nSlots;         // numAlleles() + 1
  const uint8_t* aNull;
  VCE_HD int numAlleles() const { return nSlots - 1; }
  VCE_HD bool alleleNonNull(int id) const {  // Variant.allele(id) != null, Variant.java:241
    if (id < -1 || id + 1 >= nSlots) return false;
    return aNull[slot0 + id + 1] == 0;
  }
};

// lexicographic next permutation with repeats (Permutation.step :62-90); returns false when exhausted
VCE_HD bool nextPermutation(int* p, int n) {
  if (n <= 1) return false;
  int j = n - 2;
  while (p[j] >= p[j + 1]) {
    if (--j < 0) return false;
  }
  int l = n - 1;
  while (p[j] >= p[l]) --l;
  int t = p[j]; p[j] = p[l]; p[l] = t;
  int k = j + 1;
  l = n - 1;
  while (k < l) { t = p[k]; p[k] = p[l]; p[l] = t; ++k; --l; }
  return true;
}
VCE_HD void sortSmall(int* a, int n) {
  for (int i = 1; i < n; ++i) {
    const int v = a[i];
    int j = i - 1;
    while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; --j; }
    a[j + 1] = v;
  }
}

template <class Emit>
VCE_HD void emitPermutations(const int* seq, int n, bool original, Emit& emit) {
  int p[VCE_MAX_PLOIDY];
  for (int i = 0; i < n; ++i) p[i] = seq[i];
  sortSmall(p, n);
  do { emit(p, n, original); } while (nextPermutation(p, n));
}

template <class Emit>
VCE_HD void altFill(const VariantGt& v, int* alleles, int H, int i, Emit& emit) {  // AltOrientor.fillAlleles :320-337
  if (i == H) {
    emitPermutations(alleles, H, false, emit);
    return;
  }
  const bool explicitHalfCall = v.alleleNonNull(-1);
  const int lim = alleles[i - 1];
  for (int j = -1; j <= lim; ++j) {
    if (j == 0 && !explicitHalfCall) continue;
    alleles[i] = j;
    altFill(v, alleles, H, i + 1, emit);
    // restore nothing: alleles[i..] are overwritten by the next iteration; emitPermutations works on a copy
  }
}

// Calls emit(ids, n, isOriginal) once per orientation, in the reference's order.
template <class Emit>
VCE_HD void orientVariant(int kind, int H, const VariantGt& v, Emit& emit) {
  int ids[VCE_MAX_PLOIDY];
  switch (kind) {
    case VCE_ORIENT_PHASED:
    case VCE_ORIENT_PHASED_INVERT:
      if (v.phased) {  // PhasedOrientor :131-147
        if (kind == VCE_ORIENT_PHASED_INVERT) {
          int i = v.gtPloidy;
          for (int k = 0; k < v.gtPloidy; ++k) ids[--i] = v.gt[k];
          emit(ids, v.gtPloidy, false);
        } else {
          emit(v.gt, v.gtPloidy, true);
        }
        return;
      }
      // unphased call: fall through to the unphased orientor (:145)
    case VCE_ORIENT_UNPHASED:  // UnphasedOrientor :74-99
      if (H == 2) {
        if (v.gt[0] != v.gt[1]) {
          ids[0] = v.gt[0]; ids[1] = v.gt[1];
          emit(ids, 2, true);
          ids[0] = v.gt[1]; ids[1] = v.gt[0];
          emit(ids, 2, false);
        } else {
          ids[0] = v.gt[0]; ids[1] = v.gt[0];
          emit(ids, 2, true);
        }
      } else {
        emitPermutations(v.gt, v.gtPloidy, false, emit);
      }
      return;
    case VCE_ORIENT_SQUASH_GT: {  // SQUASH_GT :166-187
      int alt[VCE_MAX_PLOIDY];
      for (int i = 0; i < v.gtPloidy; ++i) alt[i] = v.gt[i];
      sortSmall(alt, v.gtPloidy);
      int prev = 0;
      for (int i = 0; i < v.gtPloidy; ++i) {
        const int id = alt[i];
        if (id > prev && v.alleleNonNull(id)) {
          ids[0] = id;
          emit(ids, 1, false);
          prev = id;
        }
      }
      return;
    }
    case VCE_ORIENT_ALLELE_GT: {  // ALLELE_GT :206-233
      const int a = v.gt[0], b = v.gt[1];
      const bool aVar = a > 0 && v.alleleNonNull(a);
      const bool bVar = b > 0 && v.alleleNonNull(b);
      const int la = aVar ? a : b;
      const int lb = bVar ? b : a;
      if (la == lb) {
        ids[0] = 0; ids[1] = la; emit(ids, 2, false);
        ids[0] = la; ids[1] = 0; emit(ids, 2, false);
        ids[0] = la; ids[1] = la; emit(ids, 2, false);
      } else {
        ids[0] = 0; ids[1] = la; emit(ids, 2, false);
        ids[0] = 0; ids[1] = lb; emit(ids, 2, false);
        ids[0] = la; ids[1] = 0; emit(ids, 2, false);
        ids[0] = lb; ids[1] = 0; emit(ids, 2, false);
        ids[0] = la; ids[1] = lb; emit(ids, 2, false);
        ids[0] = lb; ids[1] = la; emit(ids, 2, false);
      }
      return;
    }
    case VCE_ORIENT_HAPLOID_POP:  // HAPLOID_POP :250-256
      for (int i = 0; i < v.numAlleles() - 1; ++i) {
        ids[0] = i + 1;
        emit(ids, 1, false);
      }
      return;
    case VCE_ORIENT_DIPLOID_POP: {  // DIPLOID_POP :274-289
      const bool explicitHalfCall = v.alleleNonNull(-1);
      for (int i = 1; i < v.numAlleles(); ++i) {
        for (int j = -1; j < i; ++j) {
          if (j == 0 && !explicitHalfCall) continue;
          ids[0] = i; ids[1] = j; emit(ids, 2, false);
          ids[0] = j; ids[1] = i; emit(ids, 2, false);
        }
        ids[0] = i; ids[1] = i; emit(ids, 2, false);
      }
      return;
    }
    case VCE_ORIENT_ALT: {  // AltOrientor :311-319
      int r[VCE_MAX_PLOIDY];
      for (int i = 1; i < v.numAlleles(); ++i) {
        r[0] = i;
        altFill(v, r, H, 1, emit);
      }
      return;
    }
    default:
      return;
  }
}

struct OrientCount {
  int n = 0;
  VCE_HD void operator()(const int*, int, bool) { ++n; }
};
struct OrientFill {
  int row;            // next row to write
  int H;
  int slot0, nSlots;
  const uint8_t* aNull;
  int32_t* oSlot;
  int8_t* oIds;
  uint8_t* oOrig;
  VCE_HD void operator()(const int* ids, int n, bool original) {
    for (int h = 0; h < 4; ++h) oIds[(size_t) row * 4 + h] = h < n ? (int8_t) ids[h] : (int8_t) -2;
    for (int h = 0; h < H; ++h) {
      const int id = h < n ? ids[h] : -2;
      int slot = -1;
      if (id >= -1 && id + 1 < nSlots && aNull[slot0 + id + 1] == 0) slot = slot0 + id + 1;
      oSlot[(size_t) row * H + h] = slot;
    }
    oOrig[row] = original ? 1 : 0;
    ++row;
  }
};

}  // namespace vce
#endif
