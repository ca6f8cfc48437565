// rvh_ensemble.cpp -- see rvh_ensemble.h.  Random numbers come from libc rand() exactly as in the reference, so that a
// member here has the parameters the reference would have given the same member (same seed, same draw order).
#include "rvh_ensemble.h"
#include <cstdlib>

double UniformRandom() { return (double)(rand()) / RAND_MAX; }
double GaussRandom() {   // Box-Muller, cosine branch only: two rand() calls per deviate
  double u1 = (double)(rand() + 1) / (RAND_MAX + 1.0);
  double u2 = (double)(rand()) / RAND_MAX;
  return sqrt(-2.0 * log(u1)) * cos(2.0 * RVH_PI * u2);
}
double SampleFromGamma(const double &shape, const double &scale) {   // Marsaglia & Tsang (2000); variable number of draws
  if (shape < 1.0) {
    double u = UniformRandom();
    return SampleFromGamma(shape + 1.0, scale) * pow(u, 1.0 / shape);
  }
  const double d = shape - 1.0 / 3.0;
  const double c = 1.0 / std::sqrt(9.0 * d);
  double x, v;
  while (true) {
    do {
      x = GaussRandom();
      v = 1.0 + c * x;
    } while (v <= 0.0);
    v = v * v * v;
    double u = UniformRandom();
    if (u < 1.0 - 0.0331 * x * x * x * x) return scale * d * v;
    if (log(u) < 0.5 * x * x + d * (1.0 - v + std::log(v))) return scale * d * v;
  }
}
double SampleFromDistribution(int distribution, const double distpar[3]) {
  double value = 0;
  if (distribution == DIST_UNIFORM) value = distpar[0] + (distpar[1] - distpar[0]) * UniformRandom();
  else if (distribution == DIST_NORMAL) value = distpar[0] + (distpar[1] * GaussRandom());
  else if (distribution == DIST_LOGNORMAL) value = exp(distpar[0] + distpar[1] * GaussRandom());
  else if (distribution == DIST_GAMMA) value = SampleFromGamma(distpar[0], distpar[1]);
  return value;
}

CEnsemble::CEnsemble(int num_members, const optStruct &Options) : _type(ENSEMBLE_NONE), _nMembers(num_members) { (void)Options; }
void CEnsemble::SetRandomSeed(unsigned int seed) { srand(seed); }
void CEnsemble::UpdateModel(CModel *pModel, optStruct &Options, int e) { (void)pModel; (void)Options; (void)e; }

CMonteCarloEnsemble::CMonteCarloEnsemble(int num_members, const optStruct &Options) : CEnsemble(num_members, Options) { _type = ENSEMBLE_MONTECARLO; }
void CMonteCarloEnsemble::Initialize(const CModel *pModel, const optStruct &Options) { (void)pModel; (void)Options; }
void CMonteCarloEnsemble::UpdateModel(CModel *pModel, optStruct &Options, int e) {   // src/ModelEnsemble.cpp:230-263
  CEnsemble::UpdateModel(pModel, Options, e);
  ExitGracefullyIf(e >= _nMembers, "CMonteCarloEnsemble::UpdateMode: invalid ensemble member index");
  _last.resize(_ParamDists.size());
  for (size_t i = 0; i < _ParamDists.size(); i++) {
    const double val = SampleFromDistribution(_ParamDists[i].distribution, _ParamDists[i].distpar);
    pModel->UpdateParameter(_ParamDists[i].param_class, _ParamDists[i].param_name, _ParamDists[i].class_group, val);
    _last[i] = val;
  }
  // the reference re-reads the .rvc here: every member starts from the same initial state (uploaded once per member by the caller)
}

CEnsemble *CreateEnsemble(CModel *pModel, const optStruct &Options) {
  CEnsemble *pE = NULL;
  if (Options.ensemble == ENSEMBLE_MONTECARLO) {
    CMonteCarloEnsemble *mc = new CMonteCarloEnsemble(pModel->num_members, Options);
    for (size_t i = 0; i < pModel->param_dists.size(); i++) mc->AddParamDist(pModel->param_dists[i]);
    pE = mc;
  } else if (Options.ensemble == ENSEMBLE_NONE) {
    pE = new CEnsemble(1, Options);
  } else {
    ExitGracefully("CreateEnsemble: this ensemble mode is not implemented in the B200 build yet");
  }
  pE->SetRandomSeed(Options.random_seed);
  return pE;
}
