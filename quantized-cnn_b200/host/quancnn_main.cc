// Command-line driver in the spirit of the reference's src/Main.cc + src/UnitTest.cc (which pick a mode by editing
// the source): the mode and paths are arguments here.
//   quancnn_b200 classify <mainDir> <clsNames> <imgLabels|-> <topk> <bmp>...      == UnitTest::UT_CaffeEvaWrapper
//   quancnn_b200 layers   <mainDir> [deviceCnt]                                   per-layer CalcFeatMap_* run vs the fused pass
// Output of `classify`: one line per image  "<file> gt=<name|-> time=<s> | idx:prob idx:prob ..."
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "CaffeEvaWrapper.h"

static int Classify(int argc, char** argv) {
  if (argc < 7) { fprintf(stderr, "usage: quancnn_b200 classify <mainDir> <clsNames> <imgLabels|-> <topk> <bmp>...\n"); return 2; }
  CaffeEvaWrapper w;
  const std::string labels = strcmp(argv[4], "-") == 0 ? "" : argv[4];
  if (!w.SetPath(argv[2], argv[3], labels) || !w.SetModel(ENUM_CaffeModel::AlexNet, ENUM_CompMethod::Aprx)) {
    fprintf(stderr, "%s\n", w.GetErrorMsg().c_str());
    return 1;
  }
  const int topk = atoi(argv[5]);
  for (int i = 6; i < argc; i++) {
    CaffeEvaRslt r;
    r.clsCntPred = topk;
    if (!w.Proc(argv[i], &r)) { fprintf(stderr, "%s\n", w.GetErrorMsg().c_str()); return 1; }
    printf("RESULT %s gt=%s time=%.6f |", argv[i], r.hasGrthClsName ? r.clsNameGrth.c_str() : "-", r.timeTotal);
    for (size_t k = 0; k < r.clsIdxLst.size(); k++) printf(" %d:%.6f", r.clsIdxLst[k], r.clsProbLst[k]);
    printf("\n");
  }
  return 0;
}

// Runs the network layer by layer through the per-layer CalcFeatMap_* members (host matrices in and out, exactly how
// the reference's executor calls them) and compares the result with the fused whole-network path.
static int Layers(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: quancnn_b200 layers <mainDir> [deviceCnt]\n"); return 2; }
  CaffeEva eva;
  eva.Init(true);
  if (argc >= 4) eva.SetDeviceCount(atoi(argv[3]));   // > 1: the fused pass is sharded over that many GPUs
  eva.SetModelName("AlexNet");
  eva.SetModelPath(std::string(argv[2]) + "/AlexNet/Bin.Files", "bvlc_alexnet_aCaF");
  if (!eva.LoadCaffePara()) return 1;
  const CaffePara& para = eva.GetCaffePara();
  const int N = 2;
  Matrix<float> img(N, para.imgChnIn, para.imgHeiIn, para.imgWidIn);
  uint32_t s = 12345;
  for (int i = 0, n = img.GetEleCnt(); i < n; i++) {  // the LCG test image of SURVEY.md 8(d)
    s = s * 1664525u + 1013904223u;
    img.GetDataPtr()[i] = static_cast<float>((s >> 8) & 0xFFFFu) / 65536.0f * 256.0f - 128.0f;
  }
  Matrix<float> probFused;
  eva.ExecForwardPass(img, &probFused);
  Matrix<float> cur(img);
  cur.Permute(0, 2, 3, 1);  // NCHW -> NHWC (reference CaffeEva.cc:225-228)
  bool firstFc = true;
  for (int l = 0; l < para.layerCnt; l++) {
    if (firstFc && para.layerInfoLst[l].type == ENUM_LyrType::FCnt) {
      cur.Permute(0, 3, 1, 2);  // NHWC -> NCHW before the first FC layer (reference CaffeEva.cc:236-238)
      cur.Resize(N, cur.GetEleCnt() / N);
      firstFc = false;
    }
    Matrix<float> nxt;
    eva.CalcFeatMap(cur, l, &nxt);
    cur = nxt;
  }
  double maxDiff = 0.0;
  for (int i = 0, n = cur.GetEleCnt(); i < n; i++)
    maxDiff = std::max(maxDiff, static_cast<double>(std::abs(cur.GetDataPtr()[i] - probFused.GetDataPtr()[i])));
  int best = 0;
  for (int c = 1; c < 1000; c++) if (cur.GetDataPtr()[best] < cur.GetDataPtr()[c]) best = c;
  printf("LAYERS n=%d argmax0=%d p=%.6f max|layerwise-fused|=%.3e\n", N, best, cur.GetDataPtr()[best], maxDiff);
  return maxDiff < 2e-5 ? 0 : 1;
}

int main(int argc, char** argv) {
  if (argc >= 2 && strcmp(argv[1], "classify") == 0) return Classify(argc, argv);
  if (argc >= 2 && strcmp(argv[1], "layers") == 0) return Layers(argc, argv);
  fprintf(stderr, "usage: quancnn_b200 classify|layers ...\n");
  return 2;
}
