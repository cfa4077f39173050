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
  FeatureRepository() : mSwap(true), mStartExt(0), mEndExt(0), mTargetKind(-1), mTrace(0), mPos(0) { memset(&mHeader, 0, sizeof(mHeader)); }

  void Init(bool swap, int extLeft, int extRight, int targetKind, int derivOrder, int *pDerivWinLen, const char *pCmnPath, const char *pCmnMask,
            const char *pCvnPath, const char *pCvnMask, const char *pCvgFile) {
    (void)pDerivWinLen; (void)pCmnPath; (void)pCvnPath; (void)derivOrder;
    mSwap = swap; mStartExt = extLeft; mEndExt = extRight;
    mTargetKind = targetKind;  // checked against every file's own kind: this reader converts nothing (CheckKind)
    if (pCmnMask || pCvnMask || pCvgFile) Error("CMEAN/VARSCALE normalisation files are not built into the B200 hot path (use a <bias>/<window> transform)");
  }
  void Trace(int t) { mTrace = t; }
  void AddFile(const std::string &entry) { mFiles.push_back(ParseEntry(entry)); }
  void AddFileList(const char *pFileName) {
    std::ifstream in(pFileName);
    if (!in.good()) Error(std::string("Cannot open script file ") + pFileName);
    std::string line;
    while (std::getline(in, line)) {
      size_t b = line.find_first_not_of(" \t\r"), e = line.find_last_not_of(" \t\r");
      if (b == std::string::npos) continue;
      AddFile(line.substr(b, e - b + 1));
    }
  }
  size_t QueueSize() const { return mFiles.size(); }
  void Rewind() { mPos = 0; }
  void MoveNext() { mPos++; }
  bool EndOfList() const { return mPos >= mFiles.size(); }
  const FileRecord &Current() const { return mFiles[mPos]; }
  const HtkHeader &CurrentHeader() const { return mHeader; }

  /// read the current file; STARTFRMEXT/ENDFRMEXT rows replicate the first/last frame
  void ReadFullMatrix(Matrix<BaseFloat> &rMatrix) {
    const FileRecord &rec = Current();
    FILE *f = fopen(rec.mPhysical.c_str(), "rb");
    if (!f) Error(std::string("Cannot open feature file: '") + rec.mPhysical + "'");
    unsigned char hb[12];
    if (fread(hb, 1, 12, f) != 12) { fclose(f); Error(std::string("Invalid HTK header in feature file: '") + rec.mPhysical + "'"); }
    HtkHeader h;
    memcpy(&h.mNSamples, hb, 4); memcpy(&h.mSamplePeriod, hb + 4, 4); memcpy(&h.mSampleSize, hb + 8, 2); memcpy(&h.mSampleKind, hb + 10, 2);
    if (mSwap) { h.mNSamples = Swap32(h.mNSamples); h.mSamplePeriod = Swap32(h.mSamplePeriod); h.mSampleSize = (int16_t)Swap16((uint16_t)h.mSampleSize); h.mSampleKind = Swap16(h.mSampleKind); }
    if (h.mSampleKind & 02000) { fclose(f); Error(std::string("Compressed (_C) HTK files are not built into the B200 hot path: ") + rec.mPhysical); }
    // the reference's header check (Features.cc:522-528) also bounds the sample period: a wrong byte order shows up here
    if (h.mSamplePeriod < 0 || h.mSamplePeriod > 100000 || h.mNSamples <= 0 || h.mSampleSize <= 0 || h.mSampleSize % 4 != 0) { fclose(f); Error(std::string("Invalid HTK header in feature file: '") + rec.mPhysical + "'"); }
    if (!KindOk(h.mSampleKind)) {
      fclose(f);
      char buf[256];
      snprintf(buf, sizeof(buf), "Cannot convert parameter kind 0%o of '%s' to TARGETKIND 0%o: kind conversions (_Z mean normalisation, _E/_0/_N energy "
               "columns, _D/_A/_T derivatives) are not built into the B200 hot path", (unsigned)h.mSampleKind, rec.mPhysical.c_str(), (unsigned)mTargetKind);
      Error(buf);
    }
    const int dim = h.mSampleSize / 4;
    int first = rec.mFirst < 0 ? 0 : rec.mFirst, last = rec.mLast < 0 ? h.mNSamples - 1 : rec.mLast;
    if (first > last || last >= h.mNSamples) { fclose(f); Error(std::string("Frame range out of file: ") + rec.mLogical); }
    const int n = last - first + 1;
    // The context rows of a frame RANGE come from the file's own neighbouring frames where it has them; only beyond the ends of the
    // FILE is the first/last frame replicated (Features.cc:1185-1191: from_frame/to_frame move outwards by min(ext, frames available)).
    const int lo = std::max(0, first - mStartExt), hi = std::min((int)h.mNSamples - 1, last + mEndExt);
    std::vector<float> buf((size_t)(hi - lo + 1) * dim);
    fseek(f, 12 + (long)lo * h.mSampleSize, SEEK_SET);
    if (fread(buf.data(), 4, buf.size(), f) != buf.size()) { fclose(f); Error(std::string("Cannot read feature file: '") + rec.mPhysical + "'"); }
    fclose(f);
    if (mSwap) {
      uint32_t *u = reinterpret_cast<uint32_t *>(buf.data());
      for (size_t i = 0; i < buf.size(); i++) u[i] = (uint32_t)Swap32((int32_t)u[i]);
    }
    rMatrix.Init(n + mStartExt + mEndExt, dim);
    for (int r = 0; r < n + mStartExt + mEndExt; r++) {
      int src = first - mStartExt + r;  // frame of the file
      src = src < lo ? lo : (src > hi ? hi : src);
      memcpy(rMatrix.pRowData(r), buf.data() + (size_t)(src - lo) * dim, sizeof(float) * dim);
    }
    mHeader = h;
    mHeader.mNSamples = n;
    if (mTrace & 1) std::cout << "[" << rec.mLogical << " " << n << "frm]" << std::flush;
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
  /// The reference converts a file's parameter kind to TARGETKIND (Features.cc:1120-1178: per-utterance mean normalisation for _Z,
  /// energy columns added or stripped, derivatives appended).  This reader converts nothing, so it accepts exactly the cases where the
  /// reference's conversion is the identity — TARGETKIND=ANON (the file's own kind), or the same base kind (or an ANON base on either
  /// side) with the same E/N/D/A/Z/0/T qualifiers — and refuses every other combination instead of training on different features.
  bool KindOk(int fileKind) const {
    const int conv = 0100 | 0200 | 0400 | 01000 | 04000 | 020000 | 0100000;
    const int anon_here = 077, anon_htk = 12;   // ReadParmKind's ANON / the value in HTK headers
    if (mTargetKind < 0 || mTargetKind == anon_here) return true;
    const int tb = mTargetKind & 077, fb = fileKind & 077;
    if (tb != anon_here && tb != anon_htk && fb != anon_htk && tb != fb) return false;
    return (mTargetKind & conv) == (fileKind & conv);
  }
  static int ReadParmKind(const char *str, bool) {
    static const char *names[] = {"WAVEFORM", "LPC", "LPREFC", "LPCEPSTRA", "LPDELCEP", "IREFC", "MFCC", "FBANK", "MELSPEC", "USER", "DISCRETE", "PLP", "ANON"};
    std::string s(str);
    size_t us = s.find('_');
    std::string base = s.substr(0, us);
    int kind = -1;
    for (int i = 0; i < 13; i++)
      if (!strcasecmp(base.c_str(), names[i])) kind = i == 12 ? 0x3F : i;
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
  int mStartExt, mEndExt, mTargetKind, mTrace;
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
