// Host-side check of the network text format (no GPU): the fast writer of tnet_base.h (snprintf %g, rows formatted on several
// threads) must produce byte for byte what the reference's `ostream << float` loop produces (KaldiLib/Matrix.tcc:522-532:
// default precision 6), and the stream-buffer number scanner must read back exactly what strtod reads.
#include <cstdio>
#include <random>
#include <sstream>

#include "tnet_base.h"

using namespace TNet;

static std::string reference_text(const Matrix<float> &m) {
  std::ostringstream out;
  out << "m " << m.Rows() << ' ' << m.Cols() << '\n';
  for (size_t i = 0; i < m.Rows(); i++) {
    for (size_t j = 0; j < m.Cols(); j++) out << m(i, j) << ' ';
    out << '\n';
  }
  return out.str();
}

int main() {
  std::mt19937 gen(7);
  std::normal_distribution<float> g(0.0f, 1.0f);
  std::uniform_int_distribution<int> e(-30, 30);
  const size_t shapes[][2] = {{1, 1}, {3, 5}, {0, 4}, {700, 400}, {2048, 135}, {33, 9001}};
  for (auto &sh : shapes) {
    Matrix<float> m(sh[0], sh[1]);
    for (size_t i = 0; i < sh[0] * sh[1]; i++) m.pData()[i] = g(gen) * std::pow(10.0f, (float)e(gen) / 3.0f);
    if (sh[0] * sh[1] > 8) {
      m.pData()[0] = 0.0f; m.pData()[1] = -0.0f; m.pData()[2] = INFINITY; m.pData()[3] = -INFINITY; m.pData()[4] = NAN;
      m.pData()[5] = 1e-45f; m.pData()[6] = 3.4028235e38f; m.pData()[7] = 123456.5f; m.pData()[8] = 1e6f;
    }
    std::ostringstream out;
    out << m;
    const std::string want = reference_text(m);
    if (out.str() != want) { fprintf(stderr, "writer differs from ostream<<float for %zux%zu\n", sh[0], sh[1]); return 1; }
    // read back: every value must equal (float)strtod(token).  inf / nan are written like the reference writes them but, like the
    // reference's `istream >> float`, not read back (checked below): the read-back copy carries finite values in their place
    std::string readable = want;
    if (sh[0] * sh[1] > 8) {
      m.pData()[2] = 2.0f; m.pData()[3] = -3.0f; m.pData()[4] = 4.0f;
      readable = reference_text(m);
    }
    const std::string want_read = readable;
    std::istringstream in(want_read + " v 3 1 2 3 <softmax> 4 4");
    Matrix<float> back;
    in >> back;
    if (back.Rows() != m.Rows() || back.Cols() != m.Cols()) { fprintf(stderr, "reader: wrong dims\n"); return 1; }
    std::istringstream tok(want_read);
    std::string t;
    tok >> t >> t >> t;  // "m rows cols"
    for (size_t i = 0; i < sh[0] * sh[1]; i++) {
      tok >> t;
      const float ref = (float)std::strtod(t.c_str(), NULL), got = back.pData()[i];
      if (!(ref == got || (std::isnan(ref) && std::isnan(got)))) { fprintf(stderr, "reader: element %zu: %s -> %g\n", i, t.c_str(), got); return 1; }
    }
    Vector<float> v;
    in >> v;                      // the stream continues right behind the matrix
    std::string tag;
    in >> tag;
    if (v.Dim() != 3 || v[2] != 3.0f || tag != "<softmax>") { fprintf(stderr, "reader: stream position lost behind the matrix\n"); return 1; }
  }
  // malformed input is an error, not a silent zero
  {
    std::istringstream in("m 2 2 1 2 x 4");
    Matrix<float> m;
    bool threw = false;
    try { in >> m; } catch (std::exception &) { threw = true; }
    if (!threw) { fprintf(stderr, "reader accepted a non-number\n"); return 1; }
  }
  {
    std::istringstream in("m 2 2 1 2 3");
    Matrix<float> m;
    bool threw = false;
    try { in >> m; } catch (std::exception &) { threw = true; }
    if (!threw) { fprintf(stderr, "reader accepted a truncated matrix\n"); return 1; }
  }
  // what the reference's reader refuses (istream >> float: no "nan"/"inf", no hexadecimal, nothing beyond the float range) is refused
  for (const char *bad : {"m 1 2 nan 1", "m 1 2 1 inf", "m 1 2 -inf 1", "m 1 2 0x10 1", "m 1 2 1e39 1", "m 1 2 1 -3.5e38"}) {
    std::istringstream in(bad);
    Matrix<float> m;
    bool threw = false;
    try { in >> m; } catch (std::exception &) { threw = true; }
    if (!threw) { fprintf(stderr, "reader accepted '%s'\n", bad); return 1; }
  }
  {
    std::istringstream in("m 1 4 1e-45 -0 3.4028235e38 1e-50");   // denormal, negative zero, FLT_MAX, underflow to zero: all fine
    Matrix<float> m;
    in >> m;
    if (m(0, 0) != 1e-45f || m(0, 2) != 3.4028235e38f || m(0, 3) != 0.0f) { fprintf(stderr, "reader: edge values\n"); return 1; }
  }
  printf("TEXT_IO_OK\n");
  return 0;
}
