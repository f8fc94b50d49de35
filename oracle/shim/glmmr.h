// oracle/shim/glmmr.h — STAND-IN for glmmrBase's <glmmr.h>: DData / DMatrix (TEST INFRASTRUCTURE).
// Reconstruction of SURVEY.md App. C.1-C.3 — UNVERIFIED against the glmmrBase sources (not available offline).
// Only the members the reference headers touch are provided (mcmldmatrix.h:19-30,46,59,61; mcmloptim.h:26;
// likelihood.h:42,157,160).
#pragma once
#include <RcppEigen.h>
#include <glmmr/maths.h>

namespace glmmr {

class DData {
public:
    Eigen::ArrayXXi cov_;
    Eigen::ArrayXd data_;
    Eigen::ArrayXd eff_range_;
    int B_ = 0;
    // state of the block selected by subdata(b)
    Eigen::ArrayXXi subcov_;
    int matstart_ = 0;
    int datastart_ = 0;
    int b_ = -1;

    DData(const Eigen::ArrayXXi& cov, const Eigen::ArrayXd& data, const Eigen::ArrayXd& eff_range)
        : cov_(cov), data_(data), eff_range_(eff_range) {
        for (int r = 0; r < cov_.rows(); r++) B_ = std::max(B_, cov_(r, 0) + 1);
    }
    int N() const { int q = 0; for (int b = 0; b < B_; b++) q += dim_of(b); return q; }
    int n_dim() const { return dim_of(b_); }
    int n_cov_pars() const {
        static const int np[15] = {0, 1, 1, 1, 2, 2, 1, 2, 2, 2, 2, 2, 2, 2, 1};
        int R = 0;
        for (int r = 0; r < cov_.rows(); r++) R = std::max(R, cov_(r, 4) + np[cov_(r, 2)]);
        return R;
    }
    void subdata(int b) {
        b_ = b;
        std::vector<int> rows;
        for (int r = 0; r < cov_.rows(); r++) if (cov_(r, 0) == b) rows.push_back(r);
        subcov_ = Eigen::ArrayXXi((int)rows.size(), 5);
        for (size_t k = 0; k < rows.size(); k++) for (int j = 0; j < 5; j++) subcov_((int)k, j) = cov_(rows[k], j);
        matstart_ = 0; datastart_ = 0;
        for (int bb = 0; bb < b; bb++) { matstart_ += dim_of(bb); datastart_ += dim_of(bb) * nvar_of(bb); }
    }
    int dim_of(int b) const { for (int r = 0; r < cov_.rows(); r++) if (cov_(r, 0) == b) return cov_(r, 1); return 0; }
    int nvar_of(int b) const { int s = 0; for (int r = 0; r < cov_.rows(); r++) if (cov_(r, 0) == b) s += cov_(r, 3); return s; }
};

class DMatrix {
public:
    DData* data_;
    Eigen::VectorXd gamma_;
    DMatrix(DData* data, const Eigen::ArrayXd& gamma) : data_(data), gamma_(gamma) {}
    DMatrix(DData* data, const Eigen::VectorXd& gamma) : data_(data), gamma_(gamma) {}
    void update_parameters(const Eigen::ArrayXd& g) { gamma_ = g; }
    void update_parameters(const Eigen::VectorXd& g) { gamma_ = g; }
    void update_parameters(const std::vector<double>& g) { gamma_ = Eigen::VectorXd((int)g.size()); for (size_t i = 0; i < g.size(); i++) gamma_((int)i) = g[i]; }

    // DSubMatrix::get_val(i, j) of the block selected in data_
    double get_val(int i, int j) const {
        const int n = data_->n_dim();
        double v = 1.0; int col0 = 0;
        for (int f = 0; f < data_->subcov_.rows(); f++) {
            const int id = data_->subcov_(f, 2), nv = data_->subcov_(f, 3), p0 = data_->subcov_(f, 4);
            double d2 = 0;
            for (int k = 0; k < nv; k++) {
                const double di = data_->data_(data_->datastart_ + (col0 + k) * n + i) - data_->data_(data_->datastart_ + (col0 + k) * n + j);
                d2 += di * di;
            }
            col0 += nv;
            const double d = std::sqrt(d2);
            switch (id) {
            case 1: v *= (d == 0 ? gamma_(p0) * gamma_(p0) : 0.0); break;                       // gr
            case 2: v *= std::exp(-d / gamma_(p0)); break;                                      // fexp0
            case 3: v *= std::pow(gamma_(p0), d); break;                                        // ar1
            case 4: v *= gamma_(p0) * std::exp(-d * d / (gamma_(p0 + 1) * gamma_(p0 + 1))); break;   // sqexp
            case 13: v *= gamma_(p0) * std::exp(-d / gamma_(p0 + 1)); break;                    // fexp
            case 14: v *= std::exp(-d * d / (gamma_(p0) * gamma_(p0))); break;                  // sqexp0
            default: throw std::runtime_error("shim: covariance function not reconstructed");
            }
        }
        return v;
    }
    // dense block, or its Cholesky factor by Cholesky–Banachiewicz (lower, or its transpose when upper)
    Eigen::MatrixXd gen_block_mat(int b, bool chol, bool upper) {
        if (data_->b_ != b) data_->subdata(b);   // MCMLDmatrix::loglik calls this from an OpenMP loop after selecting b itself: no write then
        const int n = data_->n_dim();
        Eigen::MatrixXd L = Eigen::MatrixXd::Zero(n, n);
        if (!chol) { for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) L(i, j) = get_val(i, j); return L; }
        for (int i = 0; i < n; i++)
            for (int j = 0; j <= i; j++) {
                double s = 0;
                for (int k = 0; k < j; k++) s += L(i, k) * L(j, k);
                L(i, j) = (i == j) ? std::sqrt(get_val(i, i) - s) : (get_val(i, j) - s) / L(j, j);
            }
        if (upper) return L.transpose();
        return L;
    }
    Eigen::MatrixXd genD(int, bool chol, bool upper) {
        const int Q = data_->N();
        Eigen::MatrixXd D = Eigen::MatrixXd::Zero(Q, Q);
        for (int b = 0; b < data_->B_; b++) {
            Eigen::MatrixXd blk = gen_block_mat(b, chol, upper);
            const int m = data_->matstart_, n = data_->n_dim();
            for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) D(m + i, m + j) = blk(i, j);
        }
        return D;
    }
};

// DSubMatrix(b, data, gamma): one block's entry function — only the reference's sparse-D variant constructs it (sparsedmatrix.h:62-66, out of scope);
// provided so that the reference's src/*.cpp compile as they are
class DSubMatrix {
public:
    DMatrix m_;
    DSubMatrix(int b, DData* data, const Eigen::ArrayXd& gamma) : m_(data, gamma) { if (data->b_ != b) data->subdata(b); }
    double get_val(int i, int j) const { return m_.get_val(i, j); }
};

}  // namespace glmmr
