// tnet_base.h — host-side containers, errors, timer and the text formats of TNet's network files.
//
// Mirrors the interface of the reference's KaldiLib pieces the GPU trainers use (reference paths under src/):
//   Matrix<T>/Vector<T> text I/O        KaldiLib/Matrix.tcc:522-600, Vector.tcc:527-571  ("m rows cols" / "v dim")
//   Error()/Warning()/KALDI_ERR          KaldiLib/Error.h:54-112
//   Timer                                KaldiLib/Timer.h:54-75
// Only the formats and the call surface are kept; the containers are plain row-major std::vector storage
// (no BLAS on the host: every contraction of the hot path runs on the GPU).
#ifndef TNETB200_BASE_H_
#define TNETB200_BASE_H_

#include <sys/time.h>

#include <cassert>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <algorithm>
#include <type_traits>
#include <limits>
#include <thread>
#include <vector>

namespace TNet {

typedef float BaseFloat;

class MyException : public std::runtime_error {
 public:
  explicit MyException(const std::string &what) : std::runtime_error(what) {}
};

inline void Error(const std::string &msg) { throw MyException("ERROR: " + msg); }
inline void Warning(const std::string &msg) { std::cerr << "WARNING: " << msg << std::endl; }
inline void TraceLog(const std::string &msg) { std::cout << "INFO: " << msg << std::endl; }

// KALDI_ERR << "text" << value;   throws at the end of the full expression
class ErrStream {
 public:
  ErrStream(const char *file, int line) { os_ << file << ":" << line << " "; }
  ~ErrStream() noexcept(false) { throw MyException("ERROR: " + os_.str()); }
  template <typename T>
  ErrStream &operator<<(const T &v) { os_ << v; return *this; }
 private:
  std::ostringstream os_;
};
#define KALDI_ERR ::TNet::ErrStream(__FILE__, __LINE__)

class Timer {
 public:
  void Start() { gettimeofday(&t0_, 0); }
  void End() { gettimeofday(&t1_, 0); }
  double Val() const { return (t1_.tv_sec - t0_.tv_sec) + 1e-6 * (t1_.tv_usec - t0_.tv_usec); }
 private:
  struct timeval t0_, t1_;
};

inline bool IsBigEndian() {
  const int a = 1;
  return *reinterpret_cast<const char *>(&a) != 1;
}

enum MatrixTransposeType { NO_TRANS, TRANS };

template <typename T>
class Vector {
 public:
  Vector() {}
  explicit Vector(size_t dim) : d_(dim, T()) {}
  size_t Dim() const { return d_.size(); }
  void Init(size_t dim) { d_.assign(dim, T()); }
  T *pData() { return d_.data(); }
  const T *pData() const { return d_.data(); }
  T &operator[](size_t i) { return d_[i]; }
  const T &operator[](size_t i) const { return d_[i]; }
  void Add(T v) { for (auto &x : d_) x += v; }
  // double accumulator, result cast to T (KaldiLib/Vector.tcc:266-278)
  T Sum() const { double s = 0.0; for (const auto &x : d_) s += x; return (T)s; }
 private:
  std::vector<T> d_;
};

template <typename T>
class Matrix {
 public:
  Matrix() : r_(0), c_(0) {}
  Matrix(size_t rows, size_t cols) : r_(0), c_(0) { Init(rows, cols); }
  Matrix(const Matrix<T> &src, MatrixTransposeType tr) : r_(0), c_(0) {
    if (tr == TRANS) {
      Init(src.Cols(), src.Rows());
      for (size_t i = 0; i < r_; i++) for (size_t j = 0; j < c_; j++) (*this)(i, j) = src(j, i);
    } else { *this = src; }
  }
  size_t Rows() const { return r_; }
  size_t Cols() const { return c_; }
  size_t Stride() const { return c_; }
  void Init(size_t rows, size_t cols, bool = true) { r_ = rows; c_ = cols; d_.assign(rows * cols, T()); }
  void Destroy() { r_ = c_ = 0; d_.clear(); }
  T *pData() { return d_.data(); }
  const T *pData() const { return d_.data(); }
  T *pRowData(size_t r) { return d_.data() + r * c_; }
  const T *pRowData(size_t r) const { return d_.data() + r * c_; }
  T &operator()(size_t r, size_t c) { return d_[r * c_ + c]; }
  const T &operator()(size_t r, size_t c) const { return d_[r * c_ + c]; }
  // NaN/Inf guard used by the trainers on every feature matrix (TNetCu.cc:386)
  void CheckData(const std::string &name = "") const {
    for (size_t i = 0; i < d_.size(); i++)
      if (std::isnan((double)d_[i]) || std::isinf((double)d_[i])) {
        std::ostringstream os;
        os << "Invalid value: " << d_[i] << " in matrix row: " << i / (c_ ? c_ : 1) << " col: " << i % (c_ ? c_ : 1) << " file: " << name;  // Matrix.h:238-251
        Error(os.str());
      }
  }
 private:
  size_t r_, c_;
  std::vector<T> d_;
};

typedef Matrix<BaseFloat> BfMatrix;
typedef Vector<BaseFloat> BfVector;

// ---- text format -------------------------------------------------------------------------------------
// Same text as the reference (KaldiLib/Matrix.tcc:522-600, Vector.tcc:527-571: `ostream << float` at the default precision of 6
// = printf's %g), produced and parsed fast enough that a 28M-weight network does not dominate a short run: numbers are
// scanned straight from the stream buffer (strtod on the tokens `istream >> float` would accept), large matrices are formatted row block by
// row block on several threads and written in order.
template <typename T>
inline bool ReadNumber(std::istream &in, T &v) {
  std::streambuf *sb = in.rdbuf();
  int ch = sb->sgetc();
  while (ch != std::char_traits<char>::eof() && (ch == ' ' || ch == '\n' || ch == '\t' || ch == '\r' || ch == '\f' || ch == '\v')) ch = sb->snextc();
  if (ch == std::char_traits<char>::eof()) { in.setstate(std::ios::eofbit | std::ios::failbit); return false; }
  char tok[64];
  int n = 0;
  while (ch != std::char_traits<char>::eof() && !(ch == ' ' || ch == '\n' || ch == '\t' || ch == '\r' || ch == '\f' || ch == '\v')) {
    if (n < 63) tok[n++] = (char)ch; else return false;
    ch = sb->snextc();
  }
  tok[n] = '\0';
  // what `istream >> float` takes, as the reference reads its files (Matrix.tcc:556-566): decimal digits, sign, point, exponent.
  // "nan" / "inf" (which `ostream << float` writes for a diverged network), hexadecimal floats and values beyond the float range
  // fail there, so they fail here: a network that the reference refuses to load is not loaded silently.
  for (int i = 0; i < n; i++) {
    const char c = tok[i];
    if (!((c >= '0' && c <= '9') || c == '+' || c == '-' || c == '.' || c == 'e' || c == 'E')) return false;
  }
  char *end = 0;
  double d = std::strtod(tok, &end);
  if (end == tok || *end != '\0') return false;
  const T f = (T)d;
  if (std::is_floating_point<T>::value && std::isinf((double)f)) return false;  // beyond the range after rounding to T
  v = f;
  return true;
}

inline bool IsTextSpace(int ch) { return ch == ' ' || ch == '\n' || ch == '\t' || ch == '\r' || ch == '\f' || ch == '\v'; }

/// one number token [b, e) by the rules of ReadNumber (the token has been cut at white space already)
template <typename T>
inline bool ParseNumberToken(const char *b, const char *e, T &v) {
  const size_t n = (size_t)(e - b);
  if (n == 0 || n > 63) return false;
  char tok[64];
  for (size_t i = 0; i < n; i++) {
    const char c = b[i];
    if (!((c >= '0' && c <= '9') || c == '+' || c == '-' || c == '.' || c == 'e' || c == 'E')) return false;
    tok[i] = c;
  }
  tok[n] = '\0';
  char *end = 0;
  const double d = std::strtod(tok, &end);
  if (end == tok || *end != '\0') return false;
  const T f = (T)d;
  if (std::is_floating_point<T>::value && std::isinf((double)f)) return false;
  v = f;
  return true;
}

/// all number tokens of the text [b, e) (which begins and ends at token boundaries); *bad = a token ReadNumber would refuse was met
/// (parsing stops there); ends[k] = offset behind token k when `ends` is given
template <typename T>
inline void ParseNumberSpan(const char *b, const char *e, std::vector<T> &out, bool *bad, std::vector<size_t> *ends) {
  const char *p = b;
  *bad = false;
  while (p < e) {
    while (p < e && IsTextSpace((unsigned char)*p)) p++;
    if (p >= e) break;
    const char *q = p;
    while (q < e && !IsTextSpace((unsigned char)*q)) q++;
    T v;
    if (!ParseNumberToken(p, q, v)) { *bad = true; return; }
    out.push_back(v);
    if (ends) ends->push_back((size_t)(q - b));
    p = q;
  }
}

/// `total` numbers from a SEEKABLE stream into p, several threads at a time: the text is taken in chunks of some megabytes that end at
/// white space, every chunk is cut into one span per thread (again at white space) and the spans are tokenised and converted in
/// parallel; the stream is left right behind the last number taken, exactly where the one-by-one reader leaves it.  Same acceptance
/// rules and the same failure as the serial reader (a 28 M-weight network: 4.5 s on one core).
template <typename T>
inline bool ReadNumbersParallel(std::istream &in, T *p, size_t total, unsigned nthr) {
  const std::streampos pos0 = in.tellg();
  if (pos0 == std::streampos(-1)) return false;   // not seekable: the caller reads one by one
  std::streambuf *sb = in.rdbuf();
  const size_t CHUNK = (size_t)32 << 20;
  std::string buf;
  size_t done = 0;
  std::streamoff consumed = 0;   // bytes of the stream behind pos0 that belong to numbers taken (and the white space before them)
  while (done < total) {
    buf.resize(CHUNK);
    const std::streamsize got = sb->sgetn(&buf[0], (std::streamsize)CHUNK);
    if (got <= 0) { in.setstate(std::ios::eofbit | std::ios::failbit); return true; }   // "true": handled here, the caller sees the failed stream
    buf.resize((size_t)got);
    if (!IsTextSpace((unsigned char)buf.back())) {   // finish the token the chunk ends in
      for (int ch = sb->sgetc(); ch != std::char_traits<char>::eof() && !IsTextSpace(ch); ch = sb->snextc()) buf.push_back((char)ch);
    }
    // spans: equal byte counts, moved forward to the next white space
    std::vector<size_t> cut(nthr + 1, buf.size());
    cut[0] = 0;
    for (unsigned t = 1; t < nthr; t++) {
      size_t c = std::max(cut[t - 1], buf.size() * t / nthr);
      while (c < buf.size() && !IsTextSpace((unsigned char)buf[c])) c++;
      cut[t] = c;
    }
    std::vector<std::vector<T> > vals(nthr);
    std::vector<char> bad(nthr, 0);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nthr; t++)
      th.emplace_back([&, t]() { bool b = false; vals[t].reserve((cut[t + 1] - cut[t]) / 6 + 16); ParseNumberSpan<T>(buf.data() + cut[t], buf.data() + cut[t + 1], vals[t], &b, NULL); bad[t] = b ? 1 : 0; });
    for (size_t t = 0; t < th.size(); t++) th[t].join();
    size_t used_bytes = buf.size();
    for (unsigned t = 0; t < nthr; t++) {
      const size_t need = total - done, have = vals[t].size();
      if (have >= need) {
        // the matrix ends inside this span: take `need` numbers and find where the last of them ends (a bad token behind them is not ours)
        std::memcpy(p + done, vals[t].data(), need * sizeof(T));
        done += need;
        std::vector<T> again; std::vector<size_t> ends; bool b2;
        again.reserve(have);
        ParseNumberSpan<T>(buf.data() + cut[t], buf.data() + cut[t + 1], again, &b2, &ends);
        used_bytes = cut[t] + ends[need - 1];
        break;
      }
      std::memcpy(p + done, vals[t].data(), have * sizeof(T));
      done += have;
      if (bad[t]) { in.setstate(std::ios::failbit); return true; }   // a token the reader refuses among the numbers it needs
    }
    consumed += (std::streamoff)used_bytes;
  }
  in.clear();
  in.seekg(pos0 + consumed);
  return true;
}

template <typename T>
std::istream &operator>>(std::istream &in, Matrix<T> &m) {
  in >> std::ws;
  if (in.peek() == 'm') {
    in.get();
    long long r = -1, c = -1;
    in >> r >> c;
    if (in.fail() || r < 0 || c < 0) throw std::runtime_error("Failed to read matrix from stream: no size\n");
    if (m.Rows() != (size_t)r || m.Cols() != (size_t)c) m.Init(r, c);
  }
  const size_t total = m.Rows() * m.Cols();
  unsigned nthr = std::thread::hardware_concurrency();
  if (nthr > 16) nthr = 16;
  if (total >= (1u << 18) && nthr >= 2 && m.Stride() == m.Cols()) {
    if (ReadNumbersParallel(in, m.pData(), total, nthr)) {
      if (in.fail()) throw std::runtime_error("Failed to read matrix from stream");
      return in;
    }
  }
  for (size_t i = 0; i < m.Rows(); i++) {
    T *p = m.pRowData(i);
    for (size_t j = 0; j < m.Cols(); j++)
      if (!ReadNumber(in, p[j])) throw std::runtime_error("Failed to read matrix from stream");
  }
  return in;
}
// rows [r0, r1) of m as text, one row per line, every number followed by a blank
template <typename T>
inline void FormatRows(const Matrix<T> &m, size_t r0, size_t r1, std::string &out) {
  out.clear();
  out.reserve((r1 - r0) * m.Cols() * 12);
  char buf[48];
  for (size_t i = r0; i < r1; i++) {
    const T *row = m.pRowData(i);
    for (size_t j = 0; j < m.Cols(); j++) {
      int n = snprintf(buf, sizeof(buf), "%g ", (double)row[j]);
      out.append(buf, (size_t)n);
    }
    out.push_back('\n');
  }
}
template <typename T>
std::ostream &operator<<(std::ostream &out, const Matrix<T> &m) {
  out << "m " << m.Rows() << ' ' << m.Cols() << '\n';
  const size_t total = m.Rows() * m.Cols();
  unsigned nthr = std::thread::hardware_concurrency();
  if (nthr > 16) nthr = 16;
  if (total < (1u << 18) || nthr < 2 || m.Rows() < 2 * nthr) {
    std::string s;
    FormatRows(m, 0, m.Rows(), s);
    out.write(s.data(), (std::streamsize)s.size());
  } else {
    // blocks of rows, `nthr` at a time, written in order (bounded memory: a block is ~rows/(4*nthr) rows of text)
    const size_t nblk = 4 * nthr, per = (m.Rows() + nblk - 1) / nblk;
    std::vector<std::string> txt(nthr);
    for (size_t b0 = 0; b0 < nblk; b0 += nthr) {
      std::vector<std::thread> th;
      for (unsigned t = 0; t < nthr && b0 + t < nblk; t++) {
        const size_t r0 = std::min(m.Rows(), (b0 + t) * per), r1 = std::min(m.Rows(), (b0 + t + 1) * per);
        th.emplace_back([&m, &txt, t, r0, r1]() { FormatRows(m, r0, r1, txt[t]); });
      }
      for (size_t t = 0; t < th.size(); t++) { th[t].join(); out.write(txt[t].data(), (std::streamsize)txt[t].size()); }
    }
  }
  if (out.fail()) throw std::runtime_error("Failed to write matrix to stream");
  return out;
}
template <typename T>
std::istream &operator>>(std::istream &in, Vector<T> &v) {
  in >> std::ws;
  if (in.peek() == 'v') {
    in.get();
    long long n = -1;
    in >> n;
    if (in.fail() || n < 0) throw std::runtime_error("Failed to read vector from stream: no size");
    if (v.Dim() != (size_t)n) v.Init(n);
  }
  for (size_t i = 0; i < v.Dim(); i++)
    if (!ReadNumber(in, v[i])) throw std::runtime_error("Failed to read vector from stream");
  return in;
}
template <typename T>
std::ostream &operator<<(std::ostream &out, const Vector<T> &v) {
  out << "v " << v.Dim() << "  ";
  for (size_t i = 0; i < v.Dim(); i++) out << v[i] << ' ';
  if (out.fail()) throw std::runtime_error("Failed to write vector to stream");
  return out;
}

}  // namespace TNet
#endif
