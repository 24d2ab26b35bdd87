// Host-side input model of the B200 engine: what parallelProgramInputs holds for the hot path
// (reference INPUTS/parallelInputs.hpp / .cpp:12-840,1066-1261,1263-1623), read from the same JSON.
// Member names follow the reference so that the propagator code reads like its constructor.
#pragma once

#include <array>
#include <complex>
#include <memory>
#include <string>
#include <vector>

#include "json.hpp"

namespace chiml_host {

typedef std::complex<double> cplx;

// UTIL/ml_consts.hpp
constexpr double SPEED_OF_LIGHT = 2.9979245e8;
constexpr double PI = 3.14159265358979323846;
inline double EPS0() { return 1 / (4 * PI * 1.0e-7 * std::pow(SPEED_OF_LIGHT, 2.0)); }
constexpr double ELEMENTARY_CHARGE = 1.6021766208000000926392586608069679194921272585907390121132132243531032145256176590919494628906250000e-19;
constexpr double HBAR = 1.0545718001391127086220667148574871851002905804237980726847871017460777904741746884127674612624536721e-34;

enum class POLARIZATION { EX, EY, EZ, HX, HY, HZ, L, R };
enum class PLSSHAPE { GAUSSIAN, BH, RECT, CONTINUOUS, RAMP_CONT, RICKER };
enum class SHAPE { SPHERE, BLOCK, CYLINDER };
enum class DIPOR { REL_TO_NORM, ISOTROPIC, UNIDIRECTIONAL, LAT_TAN, LONG_TAN };
enum class DTCCLASS { COUT, TXT, BIN, BMP, FREQ };
// DTCTYPE with the reference's numbering (UTIL/enum.hpp:17)
enum class DTCTYPE { EX, EY, EZ, HX, HY, HZ, DX, DY, DZ, BX, BY, BZ, EPOW, HPOW, PX, PY, PZ };

// one Lorentz pole (OBJECTS/Obj.hpp:18-39)
struct LorenzDipoleOscillator
{
    DIPOR dipOrE_ = DIPOR::ISOTROPIC, dipOrM_ = DIPOR::ISOTROPIC;
    double sigP_ = 0.0, sigM_ = 0.0, tau_ = 0.0, gam_ = 0.0, omg_ = 0.0;
    std::array<double, 3> uVecDipE_ = {{1.0, 1.0, 1.0}};
    // REL_TO_NORM: weights of the surface normal and of the two tangents (INPUTS/parallelInputs.cpp:1329-1357)
    double normCompWeightE_ = 0.0, tangentLatCompWeightE_ = 0.0, tangentLongCompWeightE_ = 0.0;
};

// geometry + material of one object (OBJECTS/Obj.hpp / Obj.cpp); only the shapes the configs use
class Obj
{
public:
    SHAPE shape_;
    bool ML_ = false;
    bool useOrientedDipols_ = false;
    double eps_infty_ = 1.0, mu_infty_ = 1.0;
    std::vector<double> geoParam_, geoParamML_;
    std::array<double, 3> location_;
    std::array<std::array<double, 3>, 3> unitVec_;
    std::array<double, 9> coordTransform_;
    std::vector<LorenzDipoleOscillator> pols_;
    std::vector<double> alpha_, xi_, gamma_;            // Obj::setUpConsts, OBJECTS/Obj.cpp:299-371
    std::vector<double> magAlpha_, magXi_, magGamma_;   // magnetic poles (sigma_m != 0)
    std::vector<double> chiAlpha_, chiXi_, chiGamma_, chiGammaPrev_;   // chiral poles (tau != 0)
    bool constsSet_ = false;
    std::vector<DIPOR> dipOr_;
    std::vector<std::array<double, 3>> dipE_;
    std::vector<double> dipNormCompE_, dipTanLatCompE_, dipTanLongCompE_;   // per electric pole (REL_TO_NORM orientations)
    bool identityAxes() const;                                       // the object's axes are the Cartesian ones (coordTransform_ = 1)
    // Obj::findGradient (OBJECTS/Obj.cpp:585-666): the outward surface normal at pt, for objects whose axes are the Cartesian ones
    std::array<double, 3> findGradient(const std::array<double, 3>& pt) const;

    Obj(SHAPE s, double eps, double mu, std::vector<LorenzDipoleOscillator> pols, bool ML, std::vector<double> geo,
        std::array<double, 3> loc, std::array<std::array<double, 3>, 3> uvec);
    void setUpConsts(double dt);
    bool isObj(const std::array<double, 3>& v, double dx, const std::vector<double>& geo) const;
    void addMLBuff(double d);
    // conservative half extents of the shape (geo) along the Cartesian axes; infinite where unbounded / not finite
    std::array<double, 3> halfExtent(const std::vector<double>& geo) const;
    bool relevant() const { return ML_ || !gamma_.empty() || !magGamma_.empty() || !chiGamma_.empty() || eps_infty_ != 1.0 || mu_infty_ != 1.0; }   // parallelFDTDField.hpp:880
};

struct EnergyLevel { std::vector<double> energyStates_, weights_; int nstates_ = 1, levDescribed_ = 1; };
struct RelaxOp { int n0, nf; double rate, dephasing; };

struct QEInput
{
    int object;                                  // index into objArr_
    std::vector<std::array<int, 2>> basis;       // (l, m)
    double density;                              // mol_den (SI, m^-3)
    std::vector<EnergyLevel> levels;
    std::vector<double> couplings;               // FDTD units
    std::vector<std::vector<double>> gam;        // explicit relaxation matrix rows (may be empty)
    std::vector<RelaxOp> relax;
    std::vector<std::array<int, 3>> locs;        // emitter nodes, global grid coordinates
    std::vector<int> dtcLevs;
    std::vector<std::string> dtcPopFiles;
    int dtcPopTimeInt = 1;
};

struct SourceInput
{
    POLARIZATION pol;
    std::vector<PLSSHAPE> shapes;
    std::vector<std::vector<cplx>> fxn;          // pulse parameters, amplitude appended last (parallelFDTDField.cpp:476)
    std::array<int, 3> loc, sz;
    double phi = 90.0;
};

struct DetectorInput
{
    DTCTYPE type; DTCCLASS cls; bool SI; std::string name;
    std::array<int, 3> loc, sz;
    double timeInt;
};

// a frequency-domain detector (dtc_class "freq", DTC/parallelDTC_FREQ.hpp): running DFT of the sampled fields over a box
struct FreqDtcInput
{
    DTCTYPE type; bool SI, outputMaps; std::string name;
    std::array<int, 3> loc, sz;
    int timeInt;                      // floor(Time_Interval / dt + 0.5) (parallelFDTDField.cpp:1728)
    std::vector<double> freqs;        // angular, FDTD units
};

struct FluxInput
{
    std::string name; std::array<int, 3> loc, sz; double weight; int timeInt; std::vector<double> freqs; bool SI = false, crossSec = false, save = false, load = false; std::string incdFile;
};

class Inputs
{
public:
    bool periodic_ = false;
    bool cplxFields_ = false;                     // Bloch-periodic run: the fields are complex (two real field sets coupled by the wrap copies)
    std::array<double, 3> k_point_ = {{0.0, 0.0, 0.0}};
    std::vector<FreqDtcInput> freqDtcs_;
    POLARIZATION pol_;
    int res_;
    double courant_, a_, tMax_, I0_;
    std::array<double, 3> size_, d_;
    double dt_;
    double pmlSigOptRat_, pmlKappaMax_, pmlAMax_, pmlMa_, pmlM_;
    std::array<int, 3> pmlThickness_;
    std::vector<std::shared_ptr<Obj>> objArr_;
    std::vector<SourceInput> sources_;
    std::vector<DetectorInput> detectors_;
    std::vector<FluxInput> fluxes_;
    std::vector<QEInput> qes_;

    // postProcessing: the input is only read for its grid, flux regions and detectors (chiml_flux); a TFSF block, which the host-side
    // setup cannot turn into a plan, is then skipped instead of refused
    explicit Inputs(const Json& IP, bool postProcessing = false);
    static int find_pt(double pt, double d) { return int(std::floor(pt / d + 0.5)); }   // parallelInputs.hpp:296
    double ev2FDTD(double eV) const { return eV / 4.135666e-15 * a_ / SPEED_OF_LIGHT; }   // parallelInputs.cpp:1181-1184

private:
    std::shared_ptr<Obj> jsonToObject(const Json& o);
    std::vector<LorenzDipoleOscillator> getMetal(const std::vector<double>& params) const;
    bool getMater(const std::string& mat, double& eps, double& mu, std::vector<LorenzDipoleOscillator>& pols) const;
};

// pulse value at time t (UTIL/PulseFxn.hpp:19-90), identical complex arithmetic
cplx pulseValue(PLSSHAPE shape, double t, const std::vector<cplx>& param);

} // namespace chiml_host
