// io.h — the file-side call surface the trainers' main() functions use: option parsing, feature lists and labels.
//
// Mirrors (formats, flag grammar and error behaviour only — reference paths under src/KaldiLib):
//   UserInterface      UserInterface.cc:121-267 (option map "-X fmt KEY", --KEY=VAL, -C config, -A, key normalisation),
//                      :466-490 (GetParam prefix stripping), :567-585 (GetBool), CheckCommandLineParamUse
//   FeatureRepository  Features.cc: SCP entries "logical=physical[first,last]", HTK 12-byte header, byte swapping
//                      (NATURALREADORDER), STARTFRMEXT / ENDFRMEXT frame replication (:776-849)
//   LabelRepository    Labels.cc:44-227 (MLF records, time -> frame rounding, output label map, one-hot rows)
// Not built (outside the hot path, reported as errors when requested): compressed (_C) / CRC HTK files, derivative
// expansion (TARGETKIND with _D/_A/_T), CMEAN/VARSCALE normalisation files.
#ifndef TNETB200_IO_H_
#define TNETB200_IO_H_

#include <stdint.h>
#include <strings.h>

#include <iomanip>
#include <map>
#include <set>

#include "tnet_base.h"

namespace TNet {

// ------------------------------------------------------------------------------------------ UserInterface
class UserInterface {
 public:
  struct ValueRecord {
    std::string mValue;
    char mOption;
    bool mRead;
  };

  static std::string NormalizeKey(const char *name) {
    std::string key;
    for (const char *p = name; *p; ++p)
      if (*p != '-' && *p != '_') key.push_back((char)toupper(*p));
    return key;
  }
  void InsertConfigParam(const char *pParamName, const char *value, int optionChar) {
    ValueRecord &r = mMap[NormalizeKey(pParamName)];
    r.mValue = value;
    r.mRead = false;
    r.mOption = (char)optionChar;
  }
  /// "KEY = VALUE" lines, '#' comments
  void ReadConfig(const char *pFileName) {
    std::ifstream in(pFileName);
    if (!in.good()) throw std::runtime_error(std::string("Cannot open input config file ") + pFileName);
    std::string line;
    int line_no = 0;
    while (std::getline(in, line)) {
      line_no++;
      size_t hash = line.find('#');
      if (hash != std::string::npos) line.erase(hash);
      size_t b = line.find_first_not_of(" \t\r");
      if (b == std::string::npos) continue;
      size_t eq = line.find('=');
      if (eq == std::string::npos) {
        std::ostringstream os;
        os << "Character '=' expected (" << pFileName << ":" << line_no << ")";
        throw std::runtime_error(os.str());
      }
      std::string key = Trim(line.substr(0, eq)), val = Trim(line.substr(eq + 1));
      if (val.size() >= 2 && (val[0] == '"' || val[0] == '\'') && val[val.size() - 1] == val[0]) val = val.substr(1, val.size() - 2);
      InsertConfigParam(key.c_str(), val.c_str(), 'C');
    }
  }
  /// pOptionMapping: " -x fmt KEY[=VAL]" entries; fmt 'n' = flag, 'r'/'l'/'o' = takes an argument ('l' accumulates a list)
  int ParseOptions(int argc, char *argv[], const char *pOptionMapping, const char *pToolName) {
    struct Opt { char fmt; std::string key, val; };
    std::map<char, Opt> opts;
    {
      std::istringstream is(pOptionMapping);
      std::string tok;
      while (is >> tok) {
        if (tok.size() != 2 || tok[0] != '-') throw std::runtime_error("Fatal: malformed option map");
        Opt o;
        std::string fmt, key;
        if (!(is >> fmt >> key)) throw std::runtime_error("Fatal: Unexpected end of optionMap string");
        o.fmt = fmt[0];
        size_t eq = key.find('=');
        o.key = key.substr(0, eq);
        o.val = eq == std::string::npos ? "TRUE" : key.substr(eq + 1);
        opts[tok[1]] = o;
      }
    }
    auto is_option = [](const char *s) { return s[0] == '-' && (isalpha((unsigned char)s[1]) || s[1] == '-'); };
    for (int i = 1; i < argc; i++) {  // -A : echo the command line
      if (!strcmp(argv[i], "--")) break;
      if (!strcmp(argv[i], "-A")) {
        for (int k = 0; k < argc; k++) {
          if (strchr(argv[k], ' ') || strchr(argv[k], '*')) std::cout << '\'' << argv[k] << '\'' << " "; else std::cout << argv[k] << " ";
        }
        std::cout << std::endl;
        break;
      }
    }
    for (int i = 1; i < argc; i++) {  // -C config files first
      if (!strcmp(argv[i], "--")) break;
      if (argv[i][0] != '-' || argv[i][1] != 'C') continue;
      if (argv[i][2] != '\0') ReadConfig(argv[i] + 2);
      else if (i + 1 < argc && !is_option(argv[i + 1])) ReadConfig(argv[++i]);
      else throw std::runtime_error("Config file name expected after option '-C'");
    }
    for (int i = 1; i < argc; i++) {  // --KEY=VAL
      if (!strcmp(argv[i], "--")) break;
      if (argv[i][0] != '-' || argv[i][1] != '-') continue;
      std::string s(argv[i] + 2);
      size_t eq = s.find('=');
      if (eq == std::string::npos) throw std::runtime_error(std::string("Character '=' expected after option '") + argv[i] + "'");
      InsertConfigParam((std::string(pToolName) + ":" + s.substr(0, eq)).c_str(), s.substr(eq + 1).c_str(), '-');
    }
    int optind = 1;
    std::map<char, bool> seen;
    for (; optind < argc && is_option(argv[optind]); optind++) {
      char opt = argv[optind][1];
      const char *optarg = argv[optind][2] != '\0' ? argv[optind] + 2 : NULL;
      if (opt == '-' && !optarg) return optind + 1;
      if (opt == '-') continue;
      if (opt == 'C') { if (!optarg) optind++; continue; }
      if (opt == 'A') continue;
      std::map<char, Opt>::iterator it = opts.find(opt);
      if (it == opts.end()) throw std::runtime_error(std::string("Invalid command line option '-") + opt + "'");
      std::string param = std::string(pToolName) + ":" + it->second.key;
      if (it->second.fmt == 'n') {
        if (optarg) throw std::runtime_error(std::string("Unexpected argument '") + optarg + "' after option '-" + opt + "'");
        InsertConfigParam(param.c_str(), it->second.val.c_str(), opt);
      } else {
        if (!optarg) {
          if (optind + 1 == argc || is_option(argv[optind + 1]))
            throw std::runtime_error(std::string("Argument 1 of option '-") + opt + "' expected");
          optarg = argv[++optind];
        }
        std::string v(optarg);
        if (it->second.fmt == 'l' && seen[opt]) v = std::string(GetStr(param.c_str(), "")) + "," + v;
        seen[opt] = true;
        InsertConfigParam(param.c_str(), v.c_str(), opt);
      }
    }
    for (int i = optind; i < argc; i++)
      if (is_option(argv[i])) throw std::runtime_error(std::string("No option expected after first non-option argument '") + argv[optind] + "'");
    return optind;
  }
  /// try the full name, then strip everything up to each ':' in turn (UserInterface.cc:466-490)
  ValueRecord *GetParam(const char *pParamName) {
    std::string name = NormalizeKeepColon(pParamName);
    const char *p = name.c_str();
    while (true) {
      std::map<std::string, ValueRecord>::iterator it = mMap.find(p);
      if (it != mMap.end()) { it->second.mRead = true; return &it->second; }
      p = strchr(p, ':');
      if (!p) return NULL;
      p++;
    }
  }
  const char *GetStr(const char *name, const char *dflt) { ValueRecord *v = GetParam(name); return v ? v->mValue.c_str() : dflt; }
  long GetInt(const char *name, long dflt) {
    ValueRecord *v = GetParam(name);
    if (!v) return dflt;
    char *end;
    long r = strtol(v->mValue.c_str(), &end, 0);
    if (v->mValue.empty() || *end) throw std::runtime_error(std::string("Integer number expected for ") + name + " but found '" + v->mValue + "'");
    return r;
  }
  float GetFlt(const char *name, float dflt) {
    ValueRecord *v = GetParam(name);
    if (!v) return dflt;
    char *end;
    double r = strtod(v->mValue.c_str(), &end);
    if (v->mValue.empty() || *end) throw std::runtime_error(std::string("Decimal number expected for ") + name + " but found '" + v->mValue + "'");
    return (float)r;
  }
  bool GetBool(const char *name, bool dflt) {
    ValueRecord *v = GetParam(name);
    if (!v) return dflt;
    const char *val = v->mValue.c_str();
    if (!strcasecmp(val, "TRUE") || !strcmp(val, "T")) return true;
    if (strcasecmp(val, "FALSE") && strcmp(val, "F"))
      throw std::runtime_error(std::string("TRUE or FALSE expected for ") + name + " but found '" + val + "'");
    return false;
  }
  /// every command line parameter nothing has read is an error (UserInterface.cc:657-667), short options included
  void CheckCommandLineParamUse() {
    for (std::map<std::string, ValueRecord>::iterator it = mMap.begin(); it != mMap.end(); ++it)
      if (!it->second.mRead && it->second.mOption != 'C') throw std::runtime_error(std::string("Unexpected command line parameter ") + it->first);
  }
  /// same layout as the reference's dump (UserInterface.cc:645-654): "# " marks a parameter nothing has read yet
  void PrintConfig(std::ostream &out) {
    out << "Configuration Parameters[" << mMap.size() << "]\n";
    for (std::map<std::string, ValueRecord>::iterator it = mMap.begin(); it != mMap.end(); ++it)
      out << (it->second.mRead ? "  " : "# ") << std::setw(35) << std::left << it->first << " = " << std::setw(30) << std::left
          << it->second.mValue << " # -" << it->second.mOption << std::endl;
  }

 private:
  static std::string Trim(const std::string &s) {
    size_t b = s.find_first_not_of(" \t\r"), e = s.find_last_not_of(" \t\r");
    return b == std::string::npos ? std::string() : s.substr(b, e - b + 1);
  }
  static std::string NormalizeKeepColon(const char *name) {
    std::string key;
    for (const char *p = name; *p; ++p)
      if (*p != '-' && *p != '_') key.push_back((char)toupper(*p));
    return key;
  }
  std::map<std::string, ValueRecord> mMap;
};

/// dir/base.ext from an input name (Common.cc:118-180): dir and ext replace the input's, "/./" keeps the tail as base
inline void MakeHtkFileName(char *pOut, const char *inFileName, const char *out_dir, const char *out_ext) {
  if (!strcmp(inFileName, "-")) { strcpy(pOut, "-"); return; }
  const char *base = strrchr(inFileName, '/');
  base = base ? base + 1 : inFileName;
  const char *bend = NULL;
  if (out_ext) bend = strrchr(base, '.');
  if (!bend) bend = base + strlen(base);
  const char *keep = strstr(inFileName, "/./");
  if (keep) base = keep + 3;
  std::string out;
  if (out_dir) {
    if (*out_dir) { out += out_dir; out += "/"; }
    out.append(base, bend - base);
  } else {
    out.append(inFileName, bend - inFileName);
  }
  if (out_ext && *out_ext) { out += "."; out += out_ext; }
  strcpy(pOut, out.c_str());
}

/// A whole HTK parameter file as a float matrix, big-endian on disk whatever NATURALREADORDER says — Matrix::LoadHTK of the reference
/// (Matrix.tcc:188-246), which the trainers use for target matrices with MLFTRANSC=FALSE (TNetCu.cc:402-407)
inline void LoadHtkMatrix(const char *pFileName, Matrix<BaseFloat> &rOut) {
  FILE *f = fopen(pFileName, "rb");
  if (!f) Error(std::string("Cannot open target matrix file: '") + pFileName + "'");
  unsigned char hb[12];
  if (fread(hb, 1, 12, f) != 12) { fclose(f); Error(std::string("Invalid HTK header in target matrix file: '") + pFileName + "'"); }
  const int32_t n = (int32_t)((uint32_t)hb[0] << 24 | (uint32_t)hb[1] << 16 | (uint32_t)hb[2] << 8 | (uint32_t)hb[3]);
  const int size = (int16_t)((uint16_t)hb[8] << 8 | (uint16_t)hb[9]);
  if (n <= 0 || size <= 0 || size % 4 != 0) { fclose(f); Error(std::string("Invalid HTK header in target matrix file: '") + pFileName + "'"); }
  const int cols = size / 4;
  rOut.Init(n, cols);
  std::vector<unsigned char> row((size_t)size);
  for (int r = 0; r < n; r++) {
    if (fread(row.data(), 1, row.size(), f) != row.size()) { fclose(f); Error(std::string("Cannot read target matrix file: '") + pFileName + "'"); }
    float *dst = rOut.pRowData(r);
    for (int c = 0; c < cols; c++) {
      const uint32_t u = (uint32_t)row[4 * c] << 24 | (uint32_t)row[4 * c + 1] << 16 | (uint32_t)row[4 * c + 2] << 8 | (uint32_t)row[4 * c + 3];
      memcpy(dst + c, &u, 4);
    }
  }
  fclose(f);
}

// ------------------------------------------------------------------------------------------ FeatureRepository
struct HtkHeader {
  int32_t mNSamples;
  int32_t mSamplePeriod;
  int16_t mSampleSize;
  uint16_t mSampleKind;
};

class FeatureRepository {
 public:
  struct FileRecord {
    std::string mLogical, mPhysical;
    int mFirst, mLast;  // -1 = whole file
    const std::string &Logical() const { return mLogical; }
    const std::string &Physical() const { return mPhysical; }
  };
  FeatureRepository() : mSwap(true), mStartExt(0), mEndExt(0), mTargetKind(12), mDerivOrder(-1), mHasCmn(false), mHasCvn(false), mHasCvg(false), mTrace(0), mPos(0) {
    memset(&mHeader, 0, sizeof(mHeader));
  }

  /// Arguments as FeatureRepository::Init of the reference (Features.h): targetKind from ReadParmKind; derivOrder < 0 = whatever the
  /// first file has; pDerivWinLen = derivOrder window lengths (DELTAWINDOW / ACCWINDOW / THIRDWINDOW or DERIVWINDOWS); the CMN / CVN
  /// path + mask pairs select one normalisation file per utterance, pCvgFile is one global variance-scale file.
  void Init(bool swap, int extLeft, int extRight, int targetKind, int derivOrder, int *pDerivWinLen, const char *pCmnPath, const char *pCmnMask,
            const char *pCvnPath, const char *pCvnMask, const char *pCvgFile) {
    mSwap = swap; mStartExt = extLeft; mEndExt = extRight;
    mTargetKind = targetKind;
    mDerivOrder = derivOrder;
    mDerivWin.clear();
    for (int i = 0; pDerivWinLen && i < derivOrder; i++) mDerivWin.push_back(pDerivWinLen[i]);
    mHasCmn = pCmnPath && pCmnMask; mHasCvn = pCvnPath && pCvnMask; mHasCvg = pCvgFile != NULL;
    mCmnPath = pCmnPath ? pCmnPath : ""; mCmnMask = pCmnMask ? pCmnMask : "";
    mCvnPath = pCvnPath ? pCvnPath : ""; mCvnMask = pCvnMask ? pCvnMask : "";
    mCvgFile = pCvgFile ? pCvgFile : "";
  }
  void Trace(int t) { mTrace = t; }
  void AddFile(const std::string &entry) { mFiles.push_back(ParseEntry(entry)); }
  /// pFileName = one script file or several separated by commas (a comma behind a backslash does not separate; blanks around a name
  /// are dropped; an empty name fails like a file that cannot be opened); an entry is a white-space separated token of the file, so a
  /// line may hold several (Features.cc:390-429, Tokenizer.cc)
  void AddFileList(const char *pFileName) {
    const std::string list(pFileName);
    size_t old_pos = 0, search = 0;
    while (old_pos != std::string::npos) {
      const size_t cur = list.find(',', search);
      if (cur != 0 && cur != std::string::npos && list[cur - 1] == '\\') { search = cur + 1; continue; }
      std::string name = list.substr(old_pos, cur == std::string::npos ? std::string::npos : cur - old_pos);
      old_pos = cur == std::string::npos ? cur : cur + 1;
      search = old_pos;
      const size_t b = name.find_first_not_of(" \t\r\n"), e = name.find_last_not_of(" \t\r\n");
      name = b == std::string::npos ? std::string() : name.substr(b, e - b + 1);
      std::ifstream in(name.c_str());
      if (name.empty() || !in.good()) Error(std::string("Cannot not open list file ") + name);
      std::string entry;
      while (in >> entry) AddFile(entry);
    }
  }
  size_t QueueSize() const { return mFiles.size(); }
  void Rewind() { mPos = 0; }
  void MoveNext() { mPos++; }
  bool EndOfList() const { return mPos >= mFiles.size(); }
  const FileRecord &Current() const { return mFiles[mPos]; }
  const HtkHeader &CurrentHeader() const { return mHeader; }

  /// Read the current script-file entry the way the reference does (Features.cc:1025-1440, restated): open, header (and the scale /
  /// bias vectors of a compressed file), resolve TARGETKIND against the file's own kind, take the requested frames plus the
  /// STARTFRMEXT / ENDFRMEXT context (file's own neighbours first, then the edge frame replicated), drop energy columns and derivative
  /// blocks the target does not want, per-utterance mean normalisation (_Z), derivatives the file lacks, then the CMN / CVN / global
  /// variance files.  All arithmetic in float and in the reference's order, so the matrices are bit-identical (differential tests
  /// against the reference's own reader: tests/test_host_cpu.py).
  void ReadFullMatrix(Matrix<BaseFloat> &rMatrix) {
    const FileRecord &rec = Current();
    FILE *f = fopen(rec.mPhysical.c_str(), "rb");
    if (!f) Error(std::string("Cannot open feature file: '") + rec.mPhysical + "'");
    unsigned char hb[12];
    if (fread(hb, 1, 12, f) != 12) { fclose(f); Error(std::string("Invalid HTK header in feature file: '") + rec.mPhysical + "'"); }
    HtkHeader h;
    memcpy(&h.mNSamples, hb, 4); memcpy(&h.mSamplePeriod, hb + 4, 4); memcpy(&h.mSampleSize, hb + 8, 2); memcpy(&h.mSampleKind, hb + 10, 2);
    if (mSwap) { h.mNSamples = Swap32(h.mNSamples); h.mSamplePeriod = Swap32(h.mSamplePeriod); h.mSampleSize = (int16_t)Swap16((uint16_t)h.mSampleSize); h.mSampleKind = Swap16(h.mSampleKind); }
    // the reference's header check (Features.cc:522-528) also bounds the sample period: a wrong byte order shows up here
    if (h.mSamplePeriod < 0 || h.mSamplePeriod > 100000 || h.mNSamples < 0 || h.mSampleSize <= 0) { fclose(f); Error(std::string("Invalid HTK header in feature file: '") + rec.mPhysical + "'"); }
    const bool comp = (h.mSampleKind & K_C) != 0;
    std::vector<float> scale, bias;   // compressed files: x = (int16 + bias) / scale per column, both vectors behind the header
    long data_off = 12;
    if (comp) {
      const size_t nc = (size_t)h.mSampleSize / 2;
      scale.resize(nc); bias.resize(nc);
      if (fread(scale.data(), 4, nc, f) != nc || fread(bias.data(), 4, nc, f) != nc) { fclose(f); Error(std::string("Cannot read feature file: '") + rec.mPhysical + "'"); }
      if (mSwap) { SwapFloats(scale); SwapFloats(bias); }
      h.mNSamples -= 4;   // the two float vectors count as four 16-bit "samples"
      data_off += (long)nc * 8;
    }
    int kind = h.mSampleKind & ~K_C;
    // ---- what the file has / what is wanted
    const int sd = (kind & K_T) ? 3 : (kind & K_A) ? 2 : (kind & K_D) ? 1 : 0;
    const int sE = (kind & K_E) != 0, s0 = (kind & K_0) != 0, sN = ((kind & K_N) != 0) * (sE + s0);
    // TARGETKIND ANON adopts the FIRST file's kind for the rest of the run, as the reference's member does (Features.cc:1135-1143)
    if (mTargetKind == K_ANON) mTargetKind = kind;
    else if ((mTargetKind & 077) == K_ANON) mTargetKind = (mTargetKind & ~077) | (kind & 077);
    const int tk = mTargetKind;
    const int tE = (tk & K_E) != 0, t0 = (tk & K_0) != 0, tN = ((tk & K_N) != 0) * (tE + t0);
    const int csize = comp ? 2 : 4;
    int coefs = (h.mSampleSize / csize + sN) / (sd + 1) - sE - s0;   // static coefficients without the energy columns
    const int src_vec = (coefs + sE + s0) * (sd + 1) - sN;
    if (src_vec * csize != h.mSampleSize) { fclose(f); Error(std::string("Invalid HTK header in feature file: '") + rec.mPhysical + "' mSampleSize do not match with parmKind"); }
    if (mDerivOrder < 0) mDerivOrder = sd;
    const int dord = mDerivOrder;
    if ((!sE && tE) || (!s0 && t0) || (sN && !tN) || (tN && !tE && !t0) || (tN && !dord) || (sN && !sd && dord) ||
        ((kind & 077) != (tk & 077) && (kind & 077) != K_ANON)) {
      fclose(f);
      Error(std::string("Cannot convert ") + ParmKind2Str(kind) + " to " + ParmKind2Str(tk));
    }
    if ((int)mDerivWin.size() < dord && dord > sd) { fclose(f); Error("TARGETKIND asks for derivatives but no window lengths were given (DELTAWINDOW / ACCWINDOW / THIRDWINDOW / DERIVWINDOWS)"); }
    // two combinations index one element BEFORE a feature row in the reference (Features.cc:1289, 1317 with trg_N > 0): refuse them
    const bool sentence_cmn = !mHasCmn && !(kind & K_Z) && (tk & K_Z);
    if (tN && (sentence_cmn || (sd == 0 && dord > 0))) {
      fclose(f);
      Error(std::string("Cannot convert ") + ParmKind2Str(kind) + " to " + ParmKind2Str(tk) + ": suppressing the absolute energy (_N) together with "
            "mean normalisation or with derivatives computed from the static coefficients is undefined in the reference (it indexes outside the feature row)");
    }
    const int lo_ord = std::min(sd, dord);
    const int trg_vec = (coefs + tE + t0) * (dord + 1) - tN;
    // ---- frames: the requested range, extended by the file's own neighbours, then by replication
    int first = rec.mFirst < 0 ? 0 : rec.mFirst, last = rec.mLast < 0 ? h.mNSamples - 1 : rec.mLast;
    int ext_left = mStartExt, ext_right = mEndExt;
    { const int i = std::min(first, mStartExt); first -= i; ext_left -= i; }
    { const int i = std::min(h.mNSamples - last - 1, mEndExt); last += i; ext_right -= i; }
    if (first > last || first >= h.mNSamples || last < 0 || last >= h.mNSamples) { fclose(f); Error(std::string("Invalid frame range for feature file: '") + rec.mPhysical + "'"); }
    const int nread = last - first + 1, tot = nread + ext_left + ext_right;
    std::vector<float> raw((size_t)nread * src_vec);
    fseek(f, data_off + (long)first * h.mSampleSize, SEEK_SET);
    bool ok;
    if (comp) {
      std::vector<int16_t> q(raw.size());
      ok = fread(q.data(), 2, q.size(), f) == q.size();
      for (size_t i = 0; ok && i < q.size(); i++) {
        const int16_t v = mSwap ? (int16_t)Swap16((uint16_t)q[i]) : q[i];
        const size_t c = i % (size_t)src_vec;
        raw[i] = ((float)v + bias[c]) / scale[c];
      }
    } else {
      ok = fread(raw.data(), 4, raw.size(), f) == raw.size();
      if (ok && mSwap) SwapFloats(raw);
    }
    fclose(f);
    if (!ok) Error(std::string("Cannot read feature file: '") + rec.mPhysical + "'");
    rMatrix.Init(tot, trg_vec);
    // ---- copy what the target keeps: block 0 = statics (+ _0, _E unless the absolute energy is suppressed), blocks 1.. = derivatives
    for (int i = 0; i < nread; i++) {
      const float *src = raw.data() + (size_t)i * src_vec;
      float *dst = rMatrix.pRowData(i + ext_left);
      memcpy(dst, src, sizeof(float) * coefs); src += coefs; dst += coefs;
      if (s0 && !sN) { if (t0 && !tN) *dst++ = *src; src++; }
      if (sE && !sN) { if (tE && !tN) *dst++ = *src; src++; }
      for (int b = 0; b < sd; b++) {
        if (b < lo_ord) { memcpy(dst, src, sizeof(float) * coefs); dst += coefs; }
        src += coefs;
        if (s0) { if (b < lo_ord && t0) *dst++ = *src; src++; }
        if (sE) { if (b < lo_ord && tE) *dst++ = *src; src++; }
      }
    }
    coefs += t0 + tE;   // from here on a block is `coefs` wide (block 0: coefs - tN)
    const size_t have = (size_t)(coefs * (1 + lo_ord) - tN);
    for (int i = 0; i < ext_left; i++) memcpy(rMatrix.pRowData(i), rMatrix.pRowData(ext_left), sizeof(float) * have);
    for (int i = tot - ext_right; i < tot; i++) memcpy(rMatrix.pRowData(i), rMatrix.pRowData(tot - ext_right - 1), sizeof(float) * have);
    // ---- sentence mean normalisation of the static block (float sums in frame order, Features.cc:1281-1300)
    if (sentence_cmn) {
      for (int j = 0; j < coefs; j++) {
        float norm = 0.0f;
        for (int i = 0; i < tot; i++) norm += rMatrix.pRowData(i)[j];
        norm /= tot;
        for (int i = 0; i < tot; i++) rMatrix.pRowData(i)[j] -= norm;
      }
    }
    // ---- derivatives the file does not carry: regression over +-winLen frames of the block below, edges clamped (Features.cc:1302-1343)
    for (int d = sd; d < dord; d++) {
      const int win = mDerivWin[d];
      float norm = 0.0f;
      for (int k = 1; k <= win; k++) norm += 2 * k * k;
      for (int i = 0; i < tot; i++) {
        for (int j = 0; j < coefs; j++) {
          const int col = d * coefs - tN + j;
          float acc = 0.0f;
          for (int k = 1; k <= win; k++)
            acc += k * (rMatrix.pRowData(i + std::min(tot - 1 - i, k))[col] - rMatrix.pRowData(i - std::min(i, k))[col]);
          rMatrix.pRowData(i)[col + coefs] = acc / norm;
        }
      }
    }
    h.mNSamples = tot;
    h.mSampleSize = (int16_t)(trg_vec * 4);
    kind = tk & ~(K_D | K_A | K_T);
    // ---- normalisation files
    if (mHasCmn) {
      std::string name;
      MaskCapture(rec.mLogical, mCmnMask, name);
      if (name.empty()) Error("CMN Matching failed");
      std::vector<float> v;
      ReadCepsNormFile(mCmnPath + "/" + name, kind & ~K_Z, 0, coefs, v);
      for (int i = 0; i < tot; i++)
        for (int j = tN; j < coefs; j++) rMatrix.pRowData(i)[j - tN] -= v[j];
    }
    kind |= dord == 3 ? (K_D | K_A | K_T) : dord == 2 ? (K_D | K_A) : dord == 1 ? K_D : 0;
    if (mHasCvn) {
      std::string name;
      MaskCapture(rec.mLogical, mCvnMask, name);
      std::vector<float> v;
      ReadCepsNormFile(mCvnPath + "/" + name, kind, 1, trg_vec, v);
      for (int i = 0; i < tot; i++)
        for (int j = tN; j < trg_vec; j++) rMatrix.pRowData(i)[j - tN] *= v[j];
    }
    if (mHasCvg) {
      std::vector<float> v;
      ReadCepsNormFile(mCvgFile, -1, 2, trg_vec, v);
      for (int i = 0; i < tot; i++)
        for (int j = tN; j < trg_vec; j++) rMatrix.pRowData(i)[j - tN] *= v[j];
    }
    h.mSampleKind = (uint16_t)kind;
    mHeader = h;
    mHeader.mNSamples = tot - mStartExt - mEndExt;   // frames without the context rows (what the trainers' progress lines count)
    if (mTrace & 1) std::cout << "[" << rec.mLogical << " " << mHeader.mNSamples << "frm]" << std::flush;
  }
  /// write an uncompressed HTK parameter file in the byte order the reader was configured with (Features.cc:485-495,1481-1550)
  bool WriteFeatureMatrix(const Matrix<BaseFloat> &rMatrix, const std::string &filename, int targetKind, int samplePeriod) {
    FILE *f = fopen(filename.c_str(), "wb");
    if (!f) { Error(std::string("Cannot create file:") + filename); return false; }
    int32_t n = (int32_t)rMatrix.Rows(), per = (int32_t)samplePeriod;
    int16_t size = (int16_t)(rMatrix.Cols() * sizeof(float));
    uint16_t kind = (uint16_t)targetKind;
    if (mSwap) { n = Swap32(n); per = Swap32(per); size = (int16_t)Swap16((uint16_t)size); kind = Swap16(kind); }
    unsigned char hb[12];
    memcpy(hb, &n, 4); memcpy(hb + 4, &per, 4); memcpy(hb + 8, &size, 2); memcpy(hb + 10, &kind, 2);
    bool ok = fwrite(hb, 1, 12, f) == 12;
    std::vector<uint32_t> row(rMatrix.Cols());
    for (size_t r = 0; ok && r < rMatrix.Rows(); r++) {
      memcpy(row.data(), rMatrix.pRowData(r), sizeof(float) * rMatrix.Cols());
      if (mSwap) for (size_t c = 0; c < row.size(); c++) row[c] = (uint32_t)Swap32((int32_t)row[c]);
      ok = fwrite(row.data(), 4, row.size(), f) == row.size();
    }
    fclose(f);
    if (!ok) Error(std::string("Cannot write to file:") + filename);
    return ok;
  }
  enum { K_E = 0100, K_N = 0200, K_D = 0400, K_A = 01000, K_C = 02000, K_Z = 04000, K_K = 010000, K_0 = 020000, K_V = 040000, K_T = 0100000, K_ANON = 12 };
  static const char *const *KindNames() {
    static const char *names[] = {"WAVEFORM", "LPC", "LPREFC", "LPCEPSTRA", "LPDELCEP", "IREFC", "MFCC", "FBANK", "MELSPEC", "USER", "DISCRETE", "PLP", "ANON"};
    return names;
  }
  /// "MFCC_E_D_A" from a kind word (base name, then the qualifiers in the reference's order E N D A C Z K 0 V T, Features.cc ParmKind2Str)
  static std::string ParmKind2Str(int kind) {
    const int base = kind & 077;
    std::string out = base <= 12 ? KindNames()[base] : "UNKNOWN";
    static const struct { int bit; const char *q; } quals[] = {{K_E, "_E"}, {K_N, "_N"}, {K_D, "_D"}, {K_A, "_A"}, {K_C, "_C"}, {K_Z, "_Z"}, {K_K, "_K"}, {K_0, "_0"}, {K_V, "_V"}, {K_T, "_T"}};
    for (size_t i = 0; i < sizeof(quals) / sizeof(quals[0]); i++)
      if (kind & quals[i].bit) out += quals[i].q;
    return out;
  }
  static int ReadParmKind(const char *str, bool) {
    static const char *names[] = {"WAVEFORM", "LPC", "LPREFC", "LPCEPSTRA", "LPDELCEP", "IREFC", "MFCC", "FBANK", "MELSPEC", "USER", "DISCRETE", "PLP", "ANON"};
    std::string s(str);
    size_t us = s.find('_');
    std::string base = s.substr(0, us);
    int kind = -1;
    for (int i = 0; i < 13; i++)
      if (!strcasecmp(base.c_str(), names[i])) kind = i;
    if (kind < 0) return -1;
    while (us != std::string::npos) {
      size_t nx = s.find('_', us + 1);
      std::string q = s.substr(us + 1, nx == std::string::npos ? std::string::npos : nx - us - 1);
      if (q == "E") kind |= 0100; else if (q == "N") kind |= 0200; else if (q == "D") kind |= 0400; else if (q == "A") kind |= 01000;
      else if (q == "C") kind |= 02000; else if (q == "Z") kind |= 04000; else if (q == "K") kind |= 010000; else if (q == "0") kind |= 020000;
      else if (q == "V") kind |= 040000; else if (q == "T") kind |= 0100000; else return -1;
      us = nx;
    }
    return kind;
  }

 private:
  void SwapFloats(std::vector<float> &v) const {
    uint32_t *u = reinterpret_cast<uint32_t *>(v.data());
    for (size_t i = 0; i < v.size(); i++) u[i] = (uint32_t)Swap32((int32_t)u[i]);
  }
  /// the characters a mask's '%' match in a name, concatenated (the reference's ProcessMask, StkMatch.cc:453-493: "*/" before a mask
  /// that does not start with '*', "/" before a name that does not start with '/'); false / empty when the mask does not match
  static bool CaptureGlob(const char *p, const char *t, std::string &cap) {
    for (; *p; p++, t++) {
      if (*p == '*') {
        while (*p == '*') p++;
        if (!*p) return true;
        for (; *t; t++) {
          const size_t keep = cap.size();
          if (CaptureGlob(p, t, cap)) return true;
          cap.resize(keep);
        }
        return false;
      }
      if (!*t) return false;
      if (*p == '[') Error("character classes ([...]) in file-name masks are not built into the B200 hot path");
      if (*p == '%') cap.push_back(*t);
      else if (*p != '?' && *p != *t) return false;
    }
    return !*t;
  }
  static bool MaskCapture(const std::string &name, const std::string &mask, std::string &cap) {
    const std::string p = (mask.empty() || mask[0] != '*') ? "*/" + mask : mask, t = (name.empty() || name[0] != '/') ? "/" + name : name;
    cap.clear();
    if (CaptureGlob(p.c_str(), t.c_str(), cap)) return true;
    cap.clear();
    return false;
  }
  /// "<CEPSNORM> <KIND>  <MEAN|VARIANCE> n  v1 .. vn" (type 0 mean, 1 variance -> 1/sqrt) or "<VARSCALE> n  v1 .. vn" (type 2 -> sqrt);
  /// the kind and the count must be what the features have at that point (Features.cc:97-179)
  static void ReadCepsNormFile(const std::string &file, int kind, int type, int n, std::vector<float> &out) {
    const char *tag = type == 0 ? "MEAN" : type == 1 ? "VARIANCE" : "VARSCALE", *what = type == 0 ? "CMN" : type == 1 ? "CVN" : "VarScale";
    FILE *fp = fopen(file.c_str(), "r");
    if (!fp) Error(std::string("Cannot open ") + what + " pFileName: '" + file + "'");
    char s1[80], s2[80];
    int cnt = 0;
    bool ok = true;
    if (type != 2) {
      ok = fscanf(fp, " <%64[^>]> <%64[^>]>", s1, s2) == 2 && !strcasecmp(s1, "CEPSNORM") && ReadParmKind(s2, false) == kind;
    }
    ok = ok && fscanf(fp, " <%64[^>]> %d", s1, &cnt) == 2 && !strcasecmp(s1, tag) && cnt == n;
    if (!ok) {
      fclose(fp);
      Error((type == 2 ? std::string("") : "<CEPSNORM> <" + ParmKind2Str(kind) + ">") + " <" + tag + " ... expected in " + what + " file " + file);
    }
    out.resize(n);
    for (int i = 0; i < n; i++) {
      if (fscanf(fp, " %f", &out[i]) != 1) {
        std::string msg;
        if (fscanf(fp, "%64s", s2) == 1) msg = std::string("Decimal number expected but '") + s2 + "' found in " + what + " file " + file;
        else if (feof(fp)) msg = std::string("Unexpected end of ") + what + " file " + file;
        else msg = std::string("Cannot read ") + what + " file " + file;
        fclose(fp);
        Error(msg);
      }
      if (type == 1) out[i] = (float)(1 / sqrt(out[i]));
      else if (type == 2) out[i] = (float)sqrt(out[i]);
    }
    if (fscanf(fp, "%64s", s2) == 1) { fclose(fp); Error(std::string("End of file expected but '") + s2 + "' found in " + what + " file " + file); }
    fclose(fp);
  }
  static int32_t Swap32(int32_t v) { uint32_t u = (uint32_t)v; return (int32_t)((u >> 24) | ((u >> 8) & 0xFF00) | ((u << 8) & 0xFF0000) | (u << 24)); }
  static uint16_t Swap16(uint16_t v) { return (uint16_t)((v >> 8) | (v << 8)); }
  static FileRecord ParseEntry(const std::string &e) {
    FileRecord r;
    r.mFirst = r.mLast = -1;
    std::string s = e;
    size_t eq = s.find('=');
    if (eq != std::string::npos) { r.mLogical = s.substr(0, eq); s = s.substr(eq + 1); }
    size_t lb = s.rfind('[');
    if (lb != std::string::npos && !s.empty() && s[s.size() - 1] == ']') {
      int a = -1, b = -1;
      if (sscanf(s.c_str() + lb, "[%d,%d]", &a, &b) == 2) { r.mFirst = a; r.mLast = b; s = s.substr(0, lb); }
    }
    r.mPhysical = s;
    if (r.mLogical.empty()) r.mLogical = s;
    return r;
  }
  bool mSwap;
  int mStartExt, mEndExt, mTargetKind, mDerivOrder;
  std::vector<int> mDerivWin;
  bool mHasCmn, mHasCvn, mHasCvg;
  std::string mCmnPath, mCmnMask, mCvnPath, mCvnMask, mCvgFile;
  int mTrace;
  std::vector<FileRecord> mFiles;
  size_t mPos;
  HtkHeader mHeader;
};

// ------------------------------------------------------------------------------------------ LabelRepository
class LabelRepository {
 public:
  LabelRepository() : mHasDir(false), mHasExt(false), mTrace(0) {}
  void Trace(int t) { mTrace = t; }
  /// pLabelDir / pLabelExt (SOURCETRANSCDIR / SOURCETRANSCEXT) build the label file name of an utterance from its logical feature
  /// name exactly as the reference does (MakeHtkFileName, Labels.cc:52); that name is then looked up among the MLF's record names
  /// with the reference's rules (MlfStream.cc:40-270): names without wildcards after their first character are hashed — an exact
  /// name, or "*" + the name's tail from one of its '/' (longest tail first) — the others ('*', '?', '%' inside) are patterns tried in
  /// file order; the first definition of a name wins.
  void Init(const char *pLabelMlfFile, const char *pOutputLabelMapFile, const char *pLabelDir, const char *pLabelExt) {
    mHasDir = pLabelDir != NULL;
    mDir = pLabelDir ? pLabelDir : "";
    mHasExt = pLabelExt != NULL;
    mExt = pLabelExt ? pLabelExt : "";
    ReadOutputLabelMap(pOutputLabelMapFile);
    std::ifstream in(pLabelMlfFile);
    if (!in.good()) Error(std::string("Cannot open Label MLF file: ") + pLabelMlfFile);
    std::string line;
    std::vector<std::string> *body = NULL;
    while (std::getline(in, line)) {
      if (!line.empty() && line[line.size() - 1] == '\r') line.erase(line.size() - 1);
      if (!body) {
        if (line.empty() || line[0] == '#') continue;
        if (line[0] == '"') {
          size_t q = line.rfind('"');
          body = Insert(line.substr(1, q > 0 ? q - 1 : std::string::npos));
        }
      } else {
        if (line == ".") { body = NULL; continue; }
        body->push_back(line);
      }
    }
  }
  size_t NOutputs() const { return mLabelMap.size(); }

  /// per-frame class ids; times are divided by sourceRate with round-half-up (Labels.cc:111-112); an unlabelled frame is an error
  void GenLabelIds(std::vector<int> &ids, size_t nFrames, size_t sourceRate, const char *pFeatureLogical) {
    if (nFrames < 1) KALDI_ERR << "Number of frames:" << nFrames << " is lower than 1!!!\n" << pFeatureLogical;
    const std::string label_file = LabelFileName(pFeatureLogical);
    const std::vector<std::string> *rec = Find(label_file);
    if (!rec) Error(std::string("Cannot open label MLF record: ") + label_file);
    ids.assign(nFrames, -1);
    size_t trunc_frames = 0;
    for (size_t l = 0; l < rec->size(); l++) {
      const std::string &line = (*rec)[l];
      if (line.empty() || line[0] == '#') continue;
      std::istringstream iss(line);
      unsigned long long beg, end;
      std::string state;
      if (!(iss >> beg)) KALDI_ERR << "Cannot parse column 1 (begin)\nline: " << line << "\nfile: " << pFeatureLogical << "\n";
      if (!(iss >> end)) KALDI_ERR << "Cannot parse column 2 (end)\nline: " << line << "\nfile: " << pFeatureLogical << "\n";
      if (!(iss >> state)) KALDI_ERR << "Cannot parse column 3 (state_tag)\nline: " << line << "\nfile: " << pFeatureLogical << "\n";
      beg = (beg + sourceRate / 2) / sourceRate;
      end = (end + sourceRate / 2) / sourceRate;
      std::map<std::string, int>::iterator it = mLabelMap.find(state);
      if (it == mLabelMap.end()) Error(std::string("Unknown state tag: '") + state + "' file:'" + pFeatureLogical);
      for (unsigned long long fr = beg; fr < end; fr++) {
        if (fr >= nFrames) { trunc_frames++; continue; }
        if (ids[fr] != -1) {
          std::ostringstream os;
          os << "Frame already assigned to other state, " << " file: " << pFeatureLogical << " frame: " << fr << " nframes: " << nFrames
             << " previously assigned to: (" << ids[fr] << ") now should be assigned to: " << state << "(" << it->second << ")\n";
          Error(os.str());
        }
        ids[fr] = it->second;
      }
    }
    // every frame must carry a target: the reference refuses an all-zero row of the desired matrix (Labels.cc:161-173; its test
    // `!sum == 1.0` is true exactly when the row sum is 0)
    for (size_t i = 0; i < nFrames; i++)
      if (ids[i] < 0) {
        std::ostringstream os;
        os << "Desired vector sum isn't 1.0, " << " file: " << label_file << " row: " << i << " nframes: " << nFrames
           << " sum: 0\n";
        Error(os.str());
      }
    if (trunc_frames > 10) {
      std::ostringstream os;
      os << "Truncated frames: " << trunc_frames << " Check sourcerate in features and validity of labels\n";
      Warning(os.str());
    }
  }
  /// dense one-hot rows, as the reference builds them on the host (Labels.cc:44-190)
  void GenDesiredMatrix(BfMatrix &rDesired, size_t nFrames, size_t sourceRate, const char *pFeatureLogical) {
    std::vector<int> ids;
    GenLabelIds(ids, nFrames, sourceRate, pFeatureLogical);
    rDesired.Init(nFrames, mLabelMap.size());
    for (size_t i = 0; i < nFrames; i++)
      if (ids[i] >= 0) rDesired(i, ids[i]) = 1.0f;
  }

 private:
  void ReadOutputLabelMap(const char *file) {
    std::ifstream in(file);
    if (!in.good()) Error(std::string("Cannot open OutputLabelMapFile: ") + file);
    std::string tag;
    int i = 0;
    while (in >> tag) {
      if (mLabelMap.find(tag) != mLabelMap.end()) Error(std::string("Duplicate tag in OutputLabelMapFile: ") + tag);
      mLabelMap[tag] = i++;
    }
    if (mLabelMap.empty()) Error(std::string("Empty OutputLabelMapFile: ") + file);
  }
  std::string LabelFileName(const char *pFeatureLogical) const {
    std::vector<char> buf(strlen(pFeatureLogical) + mDir.size() + mExt.size() + 8);
    MakeHtkFileName(&buf[0], pFeatureLogical, mHasDir ? mDir.c_str() : NULL, mHasExt ? mExt.c_str() : NULL);
    return std::string(&buf[0]);
  }
  struct Record { std::vector<std::string> mLines; size_t mListLimit; };
  static size_t DirDepth(const std::string &p) { size_t d = 0; for (size_t i = 0; i < p.size(); i++) d += (p[i] == '/' || p[i] == '\\'); return d; }
  /// glob match of the reference's ProcessMask (StkMatch.cc:453-493): "*/" is put before a pattern that does not start with '*', "/"
  /// before a name that does not start with '/'; '*' = any run of characters, '?' and '%' = any one character
  static bool Glob(const char *p, const char *t) {
    for (; *p; p++, t++) {
      if (*p == '*') {
        while (*p == '*') p++;
        if (!*p) return true;
        for (; *t; t++) if (Glob(p, t)) return true;
        return false;
      }
      if (!*t) return false;
      if (*p == '[') Error("character classes ([...]) in MLF record names are not built into the B200 hot path");
      if (*p != '?' && *p != '%' && *p != *t) return false;
    }
    return !*t;
  }
  static bool MaskMatch(const std::string &name, const std::string &pattern) {
    const std::string p = pattern[0] != '*' ? "*/" + pattern : pattern, t = name[0] != '/' ? "/" + name : name;
    return Glob(p.c_str(), t.c_str());
  }
  const Record *FindInHash(const std::string &name) const {
    std::map<std::string, Record>::const_iterator it;
    // deepest stored depth first: the exact name (depth "infinity"), then '*' + the tail of the name from its d-th '/' from the right
    for (std::set<size_t>::const_reverse_iterator d = mDepths.rbegin(); d != mDepths.rend(); ++d) {
      if (*d == (size_t)-1) {
        if ((it = mHash.find(name)) != mHash.end()) return &it->second;
        continue;
      }
      size_t pos = name.size();
      bool ok = true;
      if (*d == 0) pos = 0;
      else
        for (size_t i = 0; i < *d && ok; i++) {
          if (pos == 0) { ok = false; break; }
          pos = name.find_last_of("/\\", pos - 1);
          if (pos == std::string::npos) ok = false;
        }
      if (!ok) continue;
      if ((it = mHash.find("*" + name.substr(pos))) != mHash.end()) return &it->second;
    }
    return NULL;
  }
  const std::vector<std::string> *Find(const std::string &name) const {
    const Record *h = FindInHash(name);
    const size_t limit = h ? h->mListLimit : mList.size();   // a hashed name yields to the patterns defined BEFORE it only
    for (size_t i = 0; i < limit; i++)
      if (MaskMatch(name, mList[i].first)) return &mList[i].second.mLines;
    return h ? &h->mLines : NULL;
  }
  /// returns where the record's lines go (a scratch record when an earlier definition already covers the name)
  std::vector<std::string> *Insert(const std::string &name) {
    if (name.empty()) Error("Empty record name in the label MLF");
    Record r;
    r.mListLimit = mList.size();
    if (name.find_first_of("*?%", 1) == std::string::npos) {
      mDepths.insert(name[0] == '*' ? DirDepth(name) : (size_t)-1);
      if (Find(name)) { mShadowed.mLines.clear(); return &mShadowed.mLines; }   // MlfStream.cc:76-86: the more general / earlier definition stays
      return &(mHash[name] = r).mLines;
    }
    mList.push_back(std::make_pair(name, r));
    return &mList.back().second.mLines;
  }
  std::map<std::string, Record> mHash;
  std::vector<std::pair<std::string, Record> > mList;
  std::set<size_t> mDepths;
  Record mShadowed;
  std::string mDir;
  bool mHasDir, mHasExt;
  std::map<std::string, int> mLabelMap;
  std::string mExt;
  int mTrace;
};

}  // namespace TNet
#endif
