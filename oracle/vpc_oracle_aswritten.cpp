// vpc_oracle_aswritten.cpp -- CPU oracle, part 3 (TEST INFRASTRUCTURE ONLY; documentation of the reference's defects).
//
// ICP.go_hell_ICP and Matrix.ComputeEvJacobi restated EXACTLY AS WRITTEN in the C#, defects included
// (vtkPointCloud/BaseClass/ICP.cs:18-181, 274-285; Matrix.cs:30-34, 571-668):
//   (i)   `(double)(1/P.Count)` is an integer division: the accumulated sum of p y^T is multiplied by 0 (ICP.cs:53)
//   (ii)  the mean outer product is ADDED instead of subtracted (:66)
//   (iii) delta[2] = A[0,0] (always 0) instead of A[0,1] (:76)
//   (iv)  ComputeEvJacobi's row/column and eigenvector rotation loops never use their loop index: they rewrite m[p,q] / m[p,p]
//         and V[p,q] / V[p,p] cols times (Matrix.cs:636-666)
//   (v)   the quaternion is eigenvector COLUMN 0 whatever the eigenvalue order (ICP.cs:276)
//   (vi)  the composition's copy loop runs i < 9 over a 3-row matrix: IndexOutOfRangeException at i = 3 (:170-174) whenever a
//         second round does not converge
// The product and the main oracle implement the INTENDED algorithm; this file exists so that the deviations are
// executable facts (tests/test_oracle_cpu.py::test_icp_as_written_*), not claims.

#include "vpc_oracle.h"

#include <cmath>
#include <stdexcept>
#include <vector>

namespace {

struct IndexOutOfRange : std::runtime_error { IndexOutOfRange() : std::runtime_error("IndexOutOfRangeException") {} };

// Matrix.cs:7-34: rows, cols, row-major mat; the indexer only range-checks the FLAT index (it is a plain array access)
struct Mat {
  int rows, cols;
  std::vector<double> mat;
  Mat(int r, int c) : rows(r), cols(c), mat((size_t)r * c, 0.0) {}
  double& operator()(int r, int c) {
    const long long k = (long long)r * cols + c;
    if (k < 0 || k >= (long long)mat.size()) throw IndexOutOfRange();
    return mat[(size_t)k];
  }
  double operator()(int r, int c) const {
    const long long k = (long long)r * cols + c;
    if (k < 0 || k >= (long long)mat.size()) throw IndexOutOfRange();
    return mat[(size_t)k];
  }
};
Mat mul(const Mat& a, const Mat& b) {            // Matrix.StupidMultiply, Matrix.cs:500-510
  Mat r(a.rows, b.cols);
  for (int i = 0; i < r.rows; i++)
    for (int j = 0; j < r.cols; j++)
      for (int k = 0; k < a.cols; k++) r(i, j) += a(i, k) * b(k, j);
  return r;
}
Mat add(const Mat& a, const Mat& b) { Mat r(a.rows, a.cols); for (int i = 0; i < r.rows; i++) for (int j = 0; j < r.cols; j++) r(i, j) = a(i, j) + b(i, j); return r; }
Mat sub(const Mat& a, const Mat& b) { Mat r(a.rows, a.cols); for (int i = 0; i < r.rows; i++) for (int j = 0; j < r.cols; j++) r(i, j) = a(i, j) - b(i, j); return r; }
Mat scale(double n, const Mat& m) { Mat r(m.rows, m.cols); for (int i = 0; i < m.rows; i++) for (int j = 0; j < m.cols; j++) r(i, j) = m(i, j) * n; return r; }
Mat transpose(const Mat& m) { Mat r(m.cols, m.rows); for (int i = 0; i < m.rows; i++) for (int j = 0; j < m.cols; j++) r(j, i) = m(i, j); return r; }

// Matrix.ComputeEvJacobi exactly as written, Matrix.cs:571-668
bool compute_ev_jacobi_as_written(Mat& m, double* dblEigenValue, Mat& V, int nMaxIt, double eps) {
  int i, j, p = 0, q = 0, l;
  double fm, cn, sn, omega, x, y, d;
  const int cols = m.cols;
  if (V.rows != m.rows) return false;
  l = 1;
  for (i = 0; i < cols; i++) {
    V(i, i) = 1.0;
    for (j = 0; j < cols; j++)
      if (i != j) V(i, j) = 0.0;
  }
  while (true) {
    fm = 0.0;
    for (i = 1; i <= cols - 1; i++)
      for (j = 0; j <= i - 1; j++) {
        d = std::fabs(m(i, j));
        if ((i != j) && (d > fm)) { fm = d; p = i; q = j; }
      }
    if (fm < eps) {
      for (i = 0; i < cols; ++i) dblEigenValue[i] = m(i, i);
      return true;
    }
    if (l > nMaxIt) return false;
    l = l + 1;
    x = -m(p, q);
    y = (m(q, q) - m(p, p)) / 2.0;
    omega = x / std::sqrt(x * x + y * y);
    if (y < 0.0) omega = -omega;
    sn = 1.0 + std::sqrt(1.0 - omega * omega);
    sn = omega / std::sqrt(2.0 * sn);
    cn = std::sqrt(1.0 - sn * sn);
    fm = m(p, p);
    m(p, p) = fm * cn * cn + m(q, q) * sn * sn + m(p, q) * omega;
    m(q, q) = fm * sn * sn + m(q, q) * cn * cn - m(p, q) * omega;
    m(p, q) = 0.0;
    m(q, p) = 0.0;
    for (j = 0; j <= cols - 1; j++)
      if ((j != p) && (j != q)) {            // the loop index is never used (the u / w lines are commented out in the C#)
        fm = m(p, q);
        m(p, q) = fm * cn + m(p, p) * sn;
        m(p, p) = -fm * sn + m(p, p) * cn;
      }
    for (i = 0; i <= cols - 1; i++)
      if ((i != p) && (i != q)) {
        fm = m(p, q);
        m(p, q) = fm * cn + m(p, p) * sn;
        m(p, p) = -fm * sn + m(p, p) * cn;
      }
    for (i = 0; i <= cols - 1; i++) {
      fm = V(p, q);
      V(p, q) = fm * cn + V(p, p) * sn;
      V(p, p) = -fm * sn + V(p, p) * cn;
    }
  }
}

struct P3 { double X, Y, Z; };

}  // namespace

extern "C" {

// Matrix.ComputeEvJacobi as written.  a: n x n row-major (destroyed), v: n x n.  Returns the C# bool (1/0).
int vpco_jacobi_eig_as_written(double* a, int n, double* eigval, double* v, int max_it, double eps) {
  Mat m(n, n), V(n, n);
  m.mat.assign(a, a + (size_t)n * n);
  const bool ok = compute_ev_jacobi_as_written(m, eigval, V, max_it, eps);
  std::copy(m.mat.begin(), m.mat.end(), a);
  std::copy(V.mat.begin(), V.mat.end(), v);
  return ok ? 1 : 0;
}

// ICP.go_hell_ICP as written.  R (9) and T (3) are in/out like the C#'s Matrix arguments.  rounds_done = completed passes of the
// do-loop body up to the SSE update; sse_trace (nullable, max_trace entries) receives d of every round; jacobi_ok_trace likewise.
// Returns VPCO_OK if the loop ended by convergence, VPCO_E_REFERENCE_THROWS if the C# throws IndexOutOfRangeException (R, T then
// hold what the C# had written before the throw); max_rounds bounds the loop for safety (the C# has no bound).
int vpco_icp_as_written(const double* model_xyz, int64_t m, const double* data_xyz, int64_t n, double e, int32_t max_rounds, double R_io[9],
                        double T_io[3], int32_t* rounds_done, double* sse_trace, int32_t* jacobi_ok_trace, int32_t max_trace) {
  if (m <= 0 || n <= 0 || !model_xyz || !data_xyz || !R_io || !T_io) return VPCO_E_BADARG;
  std::vector<P3> model((size_t)m), data((size_t)n);
  for (int64_t i = 0; i < m; ++i) model[i] = {model_xyz[i], model_xyz[m + i], model_xyz[2 * m + i]};
  for (int64_t i = 0; i < n; ++i) data[i] = {data_xyz[i], data_xyz[n + i], data_xyz[2 * n + i]};
  Mat R(3, 3), T(3, 1);
  R.mat.assign(R_io, R_io + 9); T.mat.assign(T_io, T_io + 3);
  auto flush = [&]() { std::copy(R.mat.begin(), R.mat.end(), R_io); std::copy(T.mat.begin(), T.mat.end(), T_io); };
  double pre_d = 0.0, d = 0.0;
  int round = 0;
  std::vector<P3> P = data, Y((size_t)n);
  if (rounds_done) *rounds_done = 0;
  try {
    do {
      pre_d = d;
      Mat R1(3, 3), T1(3, 1);
      for (int64_t i = 0; i < n; i++) {                                   // FindClosestPointSet, ICP.cs:224-250
        int64_t j = 0, order = 0;
        double mn = (P[i].X - model[j].X) * (P[i].X - model[j].X) + (P[i].Y - model[j].Y) * (P[i].Y - model[j].Y) + (P[i].Z - model[j].Z) * (P[i].Z - model[j].Z);
        j++;
        for (; j < m; j++) {
          const double dd = (P[i].X - model[j].X) * (P[i].X - model[j].X) + (P[i].Y - model[j].Y) * (P[i].Y - model[j].Y) + (P[i].Z - model[j].Z) * (P[i].Z - model[j].Z);
          if (dd < mn) { mn = dd; order = j; }
        }
        Y[i] = model[order];
      }
      P3 mp{0, 0, 0}, my{0, 0, 0};                                        // CalculateMeanPoint3D, :255-273
      for (int64_t i = 0; i < n; i++) { mp.X += P[i].X; mp.Y += P[i].Y; mp.Z += P[i].Z; }
      mp.X = mp.X / n; mp.Y = mp.Y / n; mp.Z = mp.Z / n;
      for (int64_t i = 0; i < n; i++) { my.X += Y[i].X; my.Y += Y[i].Y; my.Z += Y[i].Z; }
      my.X = my.X / n; my.Y = my.Y / n; my.Z = my.Z / n;
      Mat A(3, 3), delta(3, 1), mm(3, 3);
      for (int64_t i = 0; i < n; i++) {                                   // :40-52
        Mat p(3, 1), y(1, 3);
        p(0, 0) = P[i].X; p(1, 0) = P[i].Y; p(2, 0) = P[i].Z;
        y(0, 0) = Y[i].X; y(0, 1) = Y[i].Y; y(0, 2) = Y[i].Z;
        mm = add(mm, mul(p, y));
      }
      mm = scale((double)(1 / n), mm);                                    // :53  integer division (defect i)
      Mat mean_P(3, 1), mean_Y(1, 3);
      mean_P(0, 0) = mp.X; mean_P(1, 0) = mp.Y; mean_P(2, 0) = mp.Z;
      mean_Y(0, 0) = my.X; mean_Y(0, 1) = my.Y; mean_Y(0, 2) = my.Z;
      mm = add(mm, mul(mean_P, mean_Y));                                  // :66  plus (defect ii)
      Mat m_T = transpose(mm);
      A = sub(mm, m_T);
      delta(0, 0) = A(1, 2); delta(1, 0) = A(2, 0); delta(2, 0) = A(0, 0);   // :74-76 (defect iii)
      double tr = 0.0;
      for (int i = 0; i < 3; i++) tr += mm(i, i);
      mm = add(mm, m_T);
      Mat I3(3, 3);
      I3(0, 0) = tr; I3(1, 1) = tr; I3(2, 2) = tr;
      mm = sub(mm, I3);
      Mat Q(4, 4);
      Q(0, 0) = tr;
      Q(0, 1) = delta(0, 0); Q(0, 2) = delta(1, 0); Q(0, 3) = delta(2, 0);
      Q(1, 0) = delta(0, 0); Q(2, 0) = delta(1, 0); Q(3, 0) = delta(2, 0);
      for (int i = 1; i <= 3; i++) { Q(i, 1) = mm(i - 1, 0); Q(i, 2) = mm(i - 1, 1); Q(i, 3) = mm(i - 1, 2); }
      double eigen[4];
      Mat qr(4, 4);
      const bool rs = compute_ev_jacobi_as_written(Q, eigen, qr, 100, 0.0001);   // :108-110 (defect iv)
      if (jacobi_ok_trace && round < max_trace) jacobi_ok_trace[round] = rs ? 1 : 0;
      const double q0 = qr(0, 0), q1 = qr(1, 0), q2 = qr(2, 0), q3 = qr(3, 0);   // CalculateRotation: column 0 (defect v), :274-285
      R1(0, 0) = q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3; R1(0, 1) = 2.0 * (q1 * q2 - q0 * q3); R1(0, 2) = 2.0 * (q1 * q3 + q0 * q2);
      R1(1, 0) = 2.0 * (q1 * q2 + q0 * q3); R1(1, 1) = q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3; R1(1, 2) = 2.0 * (q2 * q3 - q0 * q1);
      R1(2, 0) = 2.0 * (q1 * q3 - q0 * q2); R1(2, 1) = 2.0 * (q2 * q3 + q0 * q1); R1(2, 2) = q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3;
      Mat qt(3, 1);
      for (int i = 0; i < 3; i++) qt(i, 0) = mean_Y(0, i);
      qt = sub(qt, mul(R1, mean_P));
      for (int i = 0; i < 3; i++) T1(i, 0) = qt(0, i);                    // :124  qt[0,i] on a 3x1: flat index i, legal
      d = 0.0;
      for (int64_t p = 0; p < n; p++)                                     // :129-133
        d += (P[p].X - Y[p].X) * (P[p].X - Y[p].X) + (P[p].Y - Y[p].Y) * (P[p].Y - Y[p].Y) + (P[p].Z - Y[p].Z) * (P[p].Z - Y[p].Z);
      if (sse_trace && round < max_trace) sse_trace[round] = d;
      round++;
      if (rounds_done) *rounds_done = round;
      if (std::fabs(d - pre_d) >= e) {                                    // :149
        if (round == 1) {
          for (int i = 0; i < 3; i++) { R(i, 0) = R1(i, 0); R(i, 1) = R1(i, 1); R(i, 2) = R1(i, 2); }
          for (int i = 0; i < 3; i++) T(i, 0) = T1(i, 0);
        } else {
          Mat tempR = mul(R1, R), tempT = mul(R1, T);
          for (int i = 0; i < 9; i++) {                                   // :170-174 (defect vi): throws at i = 3
            const double a0 = tempR(i, 0); R(i, 0) = a0;
            const double a1 = tempR(i, 1); R(i, 1) = a1;
            const double a2 = tempR(i, 2); R(i, 2) = a2;
          }
          for (int i = 0; i < 3; i++) T(i, 0) = tempT(i, 0) + T1(i, 0);
        }
        for (int64_t i = 0; i < n; i++) {                                 // TransPoint(data, R, T), :195-219
          Mat p(3, 1);
          p(0, 0) = data[i].X; p(1, 0) = data[i].Y; p(2, 0) = data[i].Z;
          const Mat z = add(mul(R, p), T);
          P[i] = {z(0, 0), z(1, 0), z(2, 0)};
        }
      }
      if (max_rounds > 0 && round >= max_rounds) break;
    } while (std::fabs(d - pre_d) >= e);
  } catch (const IndexOutOfRange&) {
    flush();
    return VPCO_E_REFERENCE_THROWS;
  }
  flush();
  return VPCO_OK;
}

}  // extern "C"
