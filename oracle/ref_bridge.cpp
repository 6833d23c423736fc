// oracle/ref_bridge.cpp — builds oracle/_ref/libggp_ref_bridge.so: the reference's own unmodified headers (through
// ref_wrappers.cpp) + the reference-side binding include/ggp_bridge.h, linked to the product library libggp_b200.so.
// This is the drop-in demonstrated end to end: std::vector<MOMAdata> built by the reference's code, flattened by the
// binding, evaluated on the GPU, results put back into the reference's own containers and rendered by the reference's
// own writers - next to the reference's CPU passes on the very same objects.  TEST INFRASTRUCTURE (tests/test_gpu_bridge.py);
// the product never links it.
#include "ref_wrappers.cpp"

#include "../include/ggp_bridge.h"

extern "C" {

void* ggp_refb_open(void* h, int device) {
    try {
        return new GgpBridge(((RefForest*)h)->cells, device);
    } catch (const std::exception& e) {
        std::cerr << "ggp_refb_open: " << e.what() << "\n";
        return nullptr;
    }
}
void ggp_refb_close(void* b) { delete (GgpBridge*)b; }

// n successive evaluations through the nlopt-signature objective of the binding (carry chained like the reference's state)
int ggp_refb_loglik_chain(void* b, const double* params, int n, double* out) {
    GgpBridge* B = (GgpBridge*)b;
    std::vector<double> g;
    try {
        for (int i = 0; i < n; ++i) out[i] = -ggp_bridge_total_likelihood(pvec(params + 11 * i), g, B);
    } catch (const std::exception&) {
        return 1;
    }
    return 0;
}

// the same evaluations as ONE batched call (scan / Hessian stencil)
int ggp_refb_loglik_batch(void* b, const double* params, int n, double* out) {
    GgpBridge* B = (GgpBridge*)b;
    try {
        const std::vector<double> r = ggp_bridge_total_likelihood_batch(pvecs(params, n), *B);
        std::memcpy(out, r.data(), sizeof(double) * (size_t)n);
    } catch (const std::exception&) {
        return 1;
    }
    return 0;
}

// predictions through the binding, read back out of the reference's own per-cell vectors
int ggp_refb_predict(void* h, void* b, const double* params, int n_seg, double* mf, double* cf, double* mb, double* cb,
                     double* mp, double* cp) {
    RefForest* f = (RefForest*)h;
    try {
        ggp_bridge_predictions(*(GgpBridge*)b, pvecs(params, n_seg));
    } catch (const std::exception& e) {
        std::cerr << "ggp_refb_predict: " << e.what() << "\n";
        return 1;
    }
    for (size_t c = 0; c < f->cells.size(); ++c) {
        const MOMAdata& m = f->cells[c];
        for (size_t t = 0; t < m.mean_forward.size(); ++t) {
            const long k = f->offset[c] + (long)t;
            store4(mf + 4 * k, m.mean_forward[t]); store16(cf + 16 * k, m.cov_forward[t]);
            store4(mb + 4 * k, m.mean_backward[t]); store16(cb + 16 * k, m.cov_backward[t]);
            store4(mp + 4 * k, m.mean_prediction[t]); store16(cp + 16 * k, m.cov_prediction[t]);
        }
    }
    return 0;
}

// the dense joints text: b == NULL -> the reference's collect_joint_distributions on the CPU, else the binding's
// (GPU records rendered by the reference's Joint_vector writer).  Returns the length; copies at most cap bytes.
long ggp_refb_joints_text(void* h, void* b, const double* params, int n_seg, double tol, int precision, char* buf, long cap) {
    RefForest* f = (RefForest*)h;
    std::ostringstream out;
    out.precision(precision);
    try {
        if (b) ggp_bridge_collect_joint_distributions(*(GgpBridge*)b, pvecs(params, n_seg), out, tol, 64);
        else collect_joint_distributions(pvecs(params, n_seg), f->cells, out, tol);
    } catch (const std::exception& e) {
        std::cerr << "ggp_refb_joints_text: " << e.what() << "\n";
        return -1;
    }
    const std::string s = out.str();
    std::memcpy(buf, s.data(), (size_t)std::min<long>(cap, (long)s.size()));
    return (long)s.size();
}

}  // extern "C"
