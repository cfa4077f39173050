// Dumps what the drop-in's front end (nnet-asr_b200/host/io.h: FeatureRepository, LabelRepository) reads from a script file +
// MLF + label map, in the call order of the trainer's main loop and in the format of oracle/ref_tools/io_dump.cc (the same dump
// through the reference's own KaldiLib readers).  tests/test_host_cpu.py compares the two byte for byte.  No GPU, no CUDA.
//
//   test_feature_io <scp> <mlf> <labelmap> <start_ext> <end_ext> <swap 0|1> <out.bin> [label_dir, e.g. "*/" as the training scripts pass with -L]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>

#include "io.h"

using namespace TNet;

static void put32(FILE *f, int v) { fwrite(&v, 4, 1, f); }

int main(int argc, char **argv) {
  // fifth mode: <tool> --fea <scp> <start_ext> <end_ext> <swap 0|1> <out.bin>  features only, with the feature-side parameters the
  // trainers read (UserInterface::GetFeatureParams) taken from the environment: FEA_TARGETKIND, FEA_DERIVWINDOWS ("2_2"), FEA_DELTAWINDOW,
  // FEA_ACCWINDOW, FEA_THIRDWINDOW, FEA_CMNDIR, FEA_CMNMASK, FEA_CVNDIR, FEA_CVNMASK, FEA_CVGFILE.  Dump: int32 n, then per entry
  // int32 rows, cols, header kind, sample period, float32 rows*cols.
  if (argc == 7 && !strcmp(argv[1], "--fea")) {
    try {
      const char *tk = getenv("FEA_TARGETKIND") ? getenv("FEA_TARGETKIND") : "ANON";
      const int kind = FeatureRepository::ReadParmKind(tk, false);
      if (kind == -1) { fprintf(stderr, "Invalid TARGETKIND\n"); return 1; }
      int order, wins[8], *pw = wins;
      if (getenv("FEA_DERIVWINDOWS")) {
        order = 0;
        std::string s(getenv("FEA_DERIVWINDOWS"));
        for (size_t p = 0; (p = s.find_first_not_of("_", p)) != std::string::npos && order < 8;) {
          size_t e = s.find('_', p);
          wins[order++] = atoi(s.substr(p, e == std::string::npos ? std::string::npos : e - p).c_str());
          p = e == std::string::npos ? s.size() : e;
        }
      } else {
        order = (kind & 0100000) ? 3 : (kind & 01000) ? 2 : (kind & 0400) ? 1 : 0;
        if (order || kind != 12) {
          wins[0] = getenv("FEA_DELTAWINDOW") ? atoi(getenv("FEA_DELTAWINDOW")) : 2;
          wins[1] = getenv("FEA_ACCWINDOW") ? atoi(getenv("FEA_ACCWINDOW")) : 2;
          wins[2] = getenv("FEA_THIRDWINDOW") ? atoi(getenv("FEA_THIRDWINDOW")) : 2;
        } else { order = -1; pw = NULL; }
      }
      std::string cmn_path, cvn_path;
      const char *cmn_mask = getenv("FEA_CMNMASK"), *cvn_mask = getenv("FEA_CVNMASK");
      if (cmn_mask && getenv("FEA_CMNDIR")) cmn_path = std::string(getenv("FEA_CMNDIR")) + "/";
      if (cvn_mask && getenv("FEA_CVNDIR")) cvn_path = std::string(getenv("FEA_CVNDIR")) + "/";
      FeatureRepository repo;
      repo.Init(atoi(argv[5]) != 0, atoi(argv[3]), atoi(argv[4]), kind, order, pw, cmn_mask ? cmn_path.c_str() : NULL, cmn_mask,
                cvn_mask ? cvn_path.c_str() : NULL, cvn_mask, getenv("FEA_CVGFILE"));
      repo.AddFileList(argv[2]);
      FILE *out = fopen(argv[6], "wb");
      if (!out) { perror("out"); return 1; }
      put32(out, (int)repo.QueueSize());
      for (repo.Rewind(); !repo.EndOfList(); repo.MoveNext()) {
        Matrix<BaseFloat> m;
        repo.ReadFullMatrix(m);
        put32(out, (int)m.Rows()); put32(out, (int)m.Cols()); put32(out, (int)(repo.CurrentHeader().mSampleKind & 0xFFFF)); put32(out, (int)repo.CurrentHeader().mSamplePeriod);
        for (size_t r = 0; r < m.Rows(); r++) fwrite(m.pRowData(r), sizeof(float), m.Cols(), out);
      }
      fclose(out);
    } catch (std::exception &e) {
      fprintf(stderr, "%s\n", e.what());
      return 1;
    }
    return 0;
  }
  // sixth mode: <tool> --readlayers <network file> <out.bin>  walks a network file the way the component factory does (tag, sizes,
  // then for <biasedlinearity> the transposed weight matrix and the bias vector through the text operators) and dumps, per affine
  // layer, int32 rows, cols, float32 matrix, int32 dim, float32 vector: the stream must stand exactly behind every matrix it has read
  if (argc == 4 && !strcmp(argv[1], "--readlayers")) {
    try {
      std::ifstream in(argv[2]);
      FILE *f = fopen(argv[3], "wb");
      std::string tag;
      while (in >> tag) {
        long a = -1, b = -1;
        in >> a >> b;
        if (in.fail()) { fprintf(stderr, "sizes behind %s\n", tag.c_str()); return 1; }
        if (tag == "<biasedlinearity>") {
          Matrix<BaseFloat> m;
          Vector<BaseFloat> v;
          in >> m;
          in >> v;
          int r = (int)m.Rows(), c = (int)m.Cols(), d = (int)v.Dim();
          if (r != a || c != b || d != a) { fprintf(stderr, "layer sizes\n"); return 1; }
          fwrite(&r, 4, 1, f); fwrite(&c, 4, 1, f);
          for (int i = 0; i < r; i++) fwrite(m.pRowData(i), sizeof(float), c, f);
          fwrite(&d, 4, 1, f);
          fwrite(v.pData(), sizeof(float), d, f);
        } else if (tag != "<sigmoid>" && tag != "<softmax>") {
          fprintf(stderr, "unexpected tag %s\n", tag.c_str());
          return 1;
        }
      }
      fclose(f);
    } catch (std::exception &e) {
      fprintf(stderr, "%s\n", e.what());
      return 1;
    }
    return 0;
  }
  // fourth mode: <tool> --readmv <text file> <out.bin>  reads a matrix then a vector with the text operators of the network files
  // (Matrix.tcc:575-600, Vector.tcc:527-547) and dumps int32 rows, cols, float32 values, int32 dim, float32 values
  if (argc == 4 && !strcmp(argv[1], "--readmv")) {
    try {
      std::ifstream in(argv[2]);
      Matrix<BaseFloat> m;
      Vector<BaseFloat> v;
      in >> m;
      in >> v;
      FILE *f = fopen(argv[3], "wb");
      int r = (int)m.Rows(), c = (int)m.Cols(), d = (int)v.Dim();
      fwrite(&r, 4, 1, f); fwrite(&c, 4, 1, f);
      for (int i = 0; i < r; i++) fwrite(m.pRowData(i), sizeof(float), c, f);
      fwrite(&d, 4, 1, f);
      if (d > 0) fwrite(v.pData(), sizeof(float), d, f);
      fclose(f);
    } catch (std::exception &e) {
      fprintf(stderr, "%s\n", e.what());
      return 1;
    }
    return 0;
  }
  // third mode: <tool> --rewrite <scp entry> <start_ext> <end_ext> <swap 0|1> <out.htk>  reads one script-file entry and writes it
  // back the way TFeaCat does (TFeaCat.cc:262 / main_TFeaCatCu.cc: WriteFeatureMatrix, USER kind, the source's sample period)
  if (argc == 7 && !strcmp(argv[1], "--rewrite")) {
    try {
      FeatureRepository repo;
      repo.Init(atoi(argv[5]) != 0, atoi(argv[3]), atoi(argv[4]), FeatureRepository::ReadParmKind(getenv("TEST_TARGETKIND") ? getenv("TEST_TARGETKIND") : "ANON", false), 0, NULL, NULL, NULL, NULL, NULL, NULL);
      const char *list = "/tmp/.io_dump_entry";
      std::string tmp = std::string(argv[6]) + ".scp";
      FILE *f = fopen(tmp.c_str(), "w");
      fprintf(f, "%s\n", argv[2]);
      fclose(f);
      (void)list;
      repo.AddFileList(tmp.c_str());
      repo.Rewind();
      Matrix<BaseFloat> m;
      repo.ReadFullMatrix(m);
      repo.WriteFeatureMatrix(m, argv[6], 9 /* PARAMKIND_USER */, repo.CurrentHeader().mSamplePeriod);
    } catch (std::exception &e) {
      fprintf(stderr, "%s\n", e.what());
      return 1;
    }
    return 0;
  }
  // second mode: <tool> --htkname <in> <dir|-> <ext|->  prints MakeHtkFileName(in, dir, ext) ("-" = NULL argument)
  if (argc == 5 && !strcmp(argv[1], "--htkname")) {
    char out[4096];
    MakeHtkFileName(out, argv[2], strcmp(argv[3], "-") ? argv[3] : NULL, strcmp(argv[4], "-") ? argv[4] : NULL);
    printf("%s\n", out);
    return 0;
  }
  if (argc != 8 && argc != 9) { fprintf(stderr, "usage: test_feature_io scp mlf labelmap start_ext end_ext swap out.bin\n"); return 2; }
  const int start_ext = atoi(argv[4]), end_ext = atoi(argv[5]);
  const bool swap = atoi(argv[6]) != 0;
  try {
    FeatureRepository feature_repo;
    LabelRepository label_repo;
    feature_repo.Init(swap, start_ext, end_ext, FeatureRepository::ReadParmKind(getenv("TEST_TARGETKIND") ? getenv("TEST_TARGETKIND") : "ANON", false), 0, NULL, NULL, NULL, NULL, NULL, NULL);
    feature_repo.AddFileList(argv[1]);
    label_repo.Init(argv[2], argv[3], argc == 9 ? argv[8] : NULL, "lab");
    FILE *out = fopen(argv[7], "wb");
    if (!out) { perror("out"); return 1; }
    put32(out, (int)feature_repo.QueueSize());
    for (feature_repo.Rewind(); !feature_repo.EndOfList(); feature_repo.MoveNext()) {
      Matrix<BaseFloat> feats;
      feature_repo.ReadFullMatrix(feats);
      const std::string logical = feature_repo.Current().Logical();
      put32(out, (int)logical.size());
      fwrite(logical.data(), 1, logical.size(), out);
      put32(out, (int)feats.Rows()); put32(out, (int)feats.Cols()); put32(out, (int)feature_repo.CurrentHeader().mSamplePeriod);
      for (size_t r = 0; r < feats.Rows(); r++) fwrite(feats.pRowData(r), sizeof(float), feats.Cols(), out);
      const int rows = (int)feats.Rows() - start_ext - end_ext;
      std::vector<int> ids;  // the int32 ids the device expands to one-hot rows (tnb_onehot): -1 = unlabelled frame
      label_repo.GenLabelIds(ids, rows, feature_repo.CurrentHeader().mSamplePeriod, logical.c_str());
      BfMatrix dense;        // and the reference's dense form must say the same
      label_repo.GenDesiredMatrix(dense, rows, feature_repo.CurrentHeader().mSamplePeriod, logical.c_str());
      put32(out, (int)ids.size());
      for (size_t r = 0; r < ids.size(); r++) {
        int id = -1;
        for (size_t c = 0; c < dense.Cols(); c++)
          if (dense(r, c) != 0.0f) id = (id == -1) ? (int)c : -2;
        if (id != ids[r]) { fprintf(stderr, "GenLabelIds and GenDesiredMatrix disagree at frame %zu\n", r); return 1; }
        put32(out, id);
      }
    }
    fclose(out);
  } catch (std::exception &e) {
    fprintf(stderr, "test_feature_io: %s\n", e.what());
    return 3;
  }
  return 0;
}
