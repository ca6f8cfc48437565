// rvh_bmi.h -- CSDMS Basic Model Interface over the GPU model, with the reference's class and factory names
// (src/Raven_BMI.h:60-155, src/BMI.h:19-69): `extern "C" CRavenBMI *bmi_model_create(); void bmi_model_destroy(CRavenBMI*)`.
// Time is in days; grid 0 = HRUs, grid 1 = subbasins; errors are thrown as std::logic_error like the reference.
#ifndef RVH_BMI_H
#define RVH_BMI_H
#include <stdexcept>
#include "rvh_solver.h"

namespace bmixx {
class Bmi {   // the subset of bmi-cxx's abstract interface that CRavenBMI implements with content (src/BMI.h)
public:
  virtual ~Bmi() {}
  virtual void Initialize(std::string config_file) = 0;
  virtual void Update() = 0;
  virtual void UpdateUntil(double time) = 0;
  virtual void Finalize() = 0;
  virtual std::string GetComponentName() = 0;
  virtual int GetInputItemCount() = 0;
  virtual int GetOutputItemCount() = 0;
  virtual std::vector<std::string> GetInputVarNames() = 0;
  virtual std::vector<std::string> GetOutputVarNames() = 0;
  virtual int GetVarGrid(std::string name) = 0;
  virtual std::string GetVarType(std::string name) = 0;
  virtual std::string GetVarUnits(std::string name) = 0;
  virtual int GetVarItemsize(std::string name) = 0;
  virtual int GetVarNbytes(std::string name) = 0;
  virtual std::string GetVarLocation(std::string name) = 0;
  virtual double GetCurrentTime() = 0;
  virtual double GetStartTime() = 0;
  virtual double GetEndTime() = 0;
  virtual std::string GetTimeUnits() = 0;
  virtual double GetTimeStep() = 0;
  virtual void GetValue(std::string name, void *dest) = 0;
  virtual void GetValueAtIndices(std::string name, void *dest, int *inds, int count) = 0;
  virtual void SetValue(std::string name, void *src) = 0;
  virtual void SetValueAtIndices(std::string name, int *inds, int count, void *src) = 0;
  virtual int GetGridRank(const int grid) = 0;
  virtual int GetGridSize(const int grid) = 0;
  virtual std::string GetGridType(const int grid) = 0;
  virtual void GetGridX(const int grid, double *x) = 0;
  virtual void GetGridY(const int grid, double *y) = 0;
  virtual void GetGridZ(const int grid, double *z) = 0;
};
}  // namespace bmixx

const int GRID_HRU = 0;
const int GRID_SUBBASIN = 1;
enum bmi_var_kind { VAR_STATE_VAR, VAR_STREAMFLOW, VAR_FORCING_FUNCTION };
struct rvn_var_data {
  bmi_var_kind type; std::string name; int sv_index; int f_index;
  rvn_var_data(bmi_var_kind t, const std::string &n) : type(t), name(n), sv_index(DOESNT_EXIST), f_index(DOESNT_EXIST) {}
};

class CRavenBMI : public bmixx::Bmi {
public:
  CRavenBMI();
  ~CRavenBMI();
  void Initialize(std::string config_file) override;
  void Update() override;
  void UpdateUntil(double time) override;
  void Finalize() override;
  std::string GetComponentName() override;
  int GetInputItemCount() override { return (int)_input_vars.size(); }
  int GetOutputItemCount() override { return (int)_output_vars.size(); }
  std::vector<std::string> GetInputVarNames() override;
  std::vector<std::string> GetOutputVarNames() override;
  int GetVarGrid(std::string name) override;
  std::string GetVarType(std::string) override { return "double"; }
  std::string GetVarUnits(std::string name) override;
  int GetVarItemsize(std::string) override { return (int)sizeof(double); }
  int GetVarNbytes(std::string name) override { return GetVarItemsize(name) * GetGridSize(GetVarGrid(name)); }
  std::string GetVarLocation(std::string) override { return "node"; }
  double GetCurrentTime() override { return tt.model_time; }
  double GetStartTime() override { return 0.0; }
  double GetEndTime() override { return Options.duration; }
  std::string GetTimeUnits() override { return "d"; }
  double GetTimeStep() override { return Options.timestep; }
  void GetValue(std::string name, void *dest) override;
  void GetValueAtIndices(std::string name, void *dest, int *inds, int count) override;
  void SetValue(std::string name, void *src) override;
  void SetValueAtIndices(std::string name, int *inds, int count, void *src) override;
  int GetGridRank(const int) override { return 1; }
  int GetGridSize(const int grid) override;
  std::string GetGridType(const int) override { return "points"; }
  void GetGridX(const int grid, double *x) override;
  void GetGridY(const int grid, double *y) override;
  void GetGridZ(const int grid, double *z) override;
private:
  void _ReadConfigFile(const std::string &config_file);
  void _CheckConfigVars(std::vector<rvn_var_data> &vars);
  const rvn_var_data &_Find(const std::string &name, bool input_first) const;
  void _Fetch(const rvn_var_data &v, std::vector<double> &out);
  CModel *pModel;
  optStruct Options;
  time_struct tt;
  CGPUSimulation *_sim;
  std::vector<rvn_var_data> _input_vars, _output_vars;
  double _duration;
  std::string _base;
};

extern "C" {
CRavenBMI *bmi_model_create();
void bmi_model_destroy(CRavenBMI *ptr);
}
#endif
