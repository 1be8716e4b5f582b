// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// extern "C" entry points around the reference's own, unmodified C++ so that
// it can be driven through ctypes (SWIG is not installed in this image).  This
// file is compiled TOGETHER WITH the reference sources where they lie under
// /root/reference (see oracle/Makefile); nothing of the reference is copied
// into this repository, and the resulting libraries only ever live in the
// git-ignored oracle/_ref/.
//
//   -DREF_COLOR unset : links dlib/crf/crfwrapper/bilateralfilter/
//                       {bilateralfilter,permutohedral}.cpp
//   -DREF_COLOR set   : links dlib/crf/crfwrapper/colorbilateralfilter/
//                       {colorbilateralfilter,permutohedral}.cpp
//
// The lattice internals (offset_, barycentric_, blur_neighbors_, M_) are
// protected members of the reference's Permutohedral class
// (permutohedral.hpp:63-75); LatticeProbe derives from it to read them.
#include <cstring>

#ifdef REF_COLOR
#include "colorbilateralfilter.hpp"
#else
#include "bilateralfilter.hpp"
#endif

namespace {
class LatticeProbe : public Permutohedral {
public:
    int vertices() const { return M_; }
    int points() const { return N_; }
    int dim() const { return d_; }
    void dump(int *offset, float *bary, int *nbr) const
    {
        const size_t cnt = (size_t)N_ * (d_ + 1);
        if (offset) std::memcpy(offset, offset_, cnt * sizeof(int));
        if (bary) std::memcpy(bary, barycentric_, cnt * sizeof(float));
        if (nbr)
            for (size_t i = 0; i < (size_t)(d_ + 1) * M_; i++) {
                nbr[2 * i + 0] = blur_neighbors_[i].n1;
                nbr[2 * i + 1] = blur_neighbors_[i].n2;
            }
    }
};
}  // namespace

extern "C" {

#ifndef REF_COLOR

void ref_bilateralfilter_batch(float *images, float *ins, float *outs, int N,
                               int K, int H, int W, float sigmargb,
                               float sigmaxy)
{
    bilateralfilter_batch(images, N * 3 * H * W, ins, N * K * H * W, outs,
                          N * K * H * W, N, K, H, W, sigmargb, sigmaxy);
}

void ref_bilateralfilter(float *image, float *in, float *out, int K, int H,
                         int W, float sigmargb, float sigmaxy)
{
    bilateralfilter(image, 3 * H * W, in, K * H * W, out, K * H * W, H, W,
                    sigmargb, sigmaxy);
}

// Builds the 5-D lattice of one image; returns M.  offset/bary: [H*W*6],
// nbr: [6*M*2] (pass NULL first to learn M).
int ref_lattice_bilateral(float *image, int H, int W, float sigmargb,
                          float sigmaxy, int *offset, float *bary, int *nbr)
{
    LatticeProbe probe;
    initializePermutohedral(image, H, W, sigmargb, sigmaxy, probe);
    probe.dump(offset, bary, nbr);
    return probe.vertices();
}

#else

void ref_colorbilateralfilter_batch(float *images, float *ins, float *outs,
                                    int N, int K, int H, int W, float sigmargb,
                                    int DIM)
{
    colorbilateralfilter_batch(images, N * 3 * H * W, ins, N * K * H * W, outs,
                               N * K * H * W, N, K, H, W, sigmargb, DIM);
}

void ref_colorbilateralfilter(float *image, float *in, float *out, int K,
                              int H, int W, float sigmargb, int DIM)
{
    colorbilateralfilter(image, DIM * H * W, in, K * H * W, out, K * H * W, H,
                         W, sigmargb, DIM);
}

int ref_lattice_color(float *image, int H, int W, float sigmargb, int DIM,
                      int *offset, float *bary, int *nbr)
{
    LatticeProbe probe;
    initializePermutohedral(image, H, W, sigmargb, DIM, probe);
    probe.dump(offset, bary, nbr);
    return probe.vertices();
}

#endif

int ref_omp_max_threads(void) { return omp_get_max_threads(); }
void ref_omp_set_threads(int n) { omp_set_num_threads(n); }

}  // extern "C"
