// JSON -> Inputs (the hot-path subset of the reference's parallelProgramInputs constructor,
// INPUTS/parallelInputs.cpp:12-840; materials :1066-1261; objects :1263-1623) and the geometry /
// material-constant code of OBJECTS/Obj.cpp that the propagator setup needs.
#include "inputs.hpp"

#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>

namespace chiml_host {

namespace {

template <typename T> std::vector<T> as_vector(const Json& pt, const std::string& key)
{
    std::vector<T> r;
    for(const auto& item : pt.child(key).kids) r.push_back(item.second.value<T>());
    return r;
}
template <typename T> std::vector<T> as_vector(const Json& pt, const std::string& key, T dflt, int sz)
{
    std::vector<T> r;
    try { for(const auto& item : pt.child(key).kids) r.push_back(item.second.value<T>()); }
    catch(std::exception&) { r = std::vector<T>(sz, dflt); }
    return r;
}
// parallelInputs.hpp:382-400: missing key or bad value -> all three entries = default
template <typename T> std::array<T, 3> as_ptArr(const Json& pt, const std::string& key, T dflt = 0)
{
    std::array<T, 3> r = {{0, 0, 0}};
    try
    {
        int ii = 0;
        for(const auto& item : pt.child(key).kids) { if(ii < 3) r[ii] = item.second.value<T>(); ++ii; }
    }
    catch(std::exception&) { r = {{dflt, dflt, dflt}}; }
    return r;
}

POLARIZATION string2pol(const std::string& p)
{
    if(p == "Ex") return POLARIZATION::EX;
    if(p == "Ey") return POLARIZATION::EY;
    if(p == "Ez") return POLARIZATION::EZ;
    if(p == "Hx") return POLARIZATION::HX;
    if(p == "Hy") return POLARIZATION::HY;
    if(p == "Hz") return POLARIZATION::HZ;
    if(p == "L") return POLARIZATION::L;
    if(p == "R") return POLARIZATION::R;
    throw std::logic_error("POLARIZATION undefined");
}
PLSSHAPE string2prof(const std::string& p)
{
    if(p == "gaussian") return PLSSHAPE::GAUSSIAN;
    if(p == "BH") return PLSSHAPE::BH;
    if(p == "rectangle") return PLSSHAPE::RECT;
    if(p == "continuous") return PLSSHAPE::CONTINUOUS;
    if(p == "ricker") return PLSSHAPE::RICKER;
    if(p == "ramped_cont") return PLSSHAPE::RAMP_CONT;
    throw std::logic_error("Pulse shape undefined");
}
DTCTYPE string2out(const std::string& t)
{
    static const char* names[] = {"Ex", "Ey", "Ez", "Hx", "Hy", "Hz", "Dx", "Dy", "Dz", "Bx", "By", "Bz", "E_pow", "H_pow", "Px", "Py", "Pz"};
    for(int i = 0; i < 17; ++i) if(t == names[i]) return DTCTYPE(i);
    throw std::logic_error("detector type " + t + " is outside the covered hot path");
}
DTCCLASS string2dtcclass(const std::string& c)
{
    if(c == "bin") return DTCCLASS::BIN;
    if(c == "bmp") return DTCCLASS::BMP;
    if(c == "txt") return DTCCLASS::TXT;
    if(c == "cout") return DTCCLASS::COUT;
    if(c == "freq") return DTCCLASS::FREQ;
    throw std::logic_error("detector class undefined");
}
DIPOR string2dipor(const std::string& s)
{
    if(s == "isotropic") return DIPOR::ISOTROPIC;
    if(s == "unidirectional") return DIPOR::UNIDIRECTIONAL;
    if(s == "normal" || s == "tangent" || s == "rel_norm") return DIPOR::REL_TO_NORM;
    if(s == "lat_tangent") return DIPOR::LAT_TAN;
    if(s == "long_tangent") return DIPOR::LONG_TAN;
    throw std::logic_error("The dipole orientation style is undefined");
}

// Rakic et al., Applied Optics 37, 5271 (1998), Lorentz-Drude parameters {omega_p, (f, Gamma, omega) ...} in eV,
// as tabulated in the reference (UTIL/dielectric_params.hpp:17-28)
const std::vector<double> AG_MAT = {9.01, 0.845, 0.048, 0.0, 0.065, 3.886, 0.816, 0.124, 0.452, 4.481, 0.011, 0.065, 8.185, 0.840, 0.916, 9.083, 5.646, 2.419, 20.29};
const std::vector<double> AU_MAT = {9.03, 0.760, 0.053, 0.0, 0.024, 0.241, 0.415, 0.010, 0.345, 0.830, 0.071, 0.870, 2.969, 0.601, 2.494, 4.304, 4.384, 2.214, 13.32};
const std::vector<double> AL_MAT = {14.98, 0.523, 0.047, 0.0, 0.227, 0.333, 0.162, 0.050, 0.312, 1.544, 0.166, 1.351, 1.808, 0.030, 3.382, 3.473};
const std::vector<double> CU_MAT = {10.83, 0.575, 0.030, 0.0, 0.061, 0.378, 0.291, 0.104, 1.056, 2.957, 0.723, 2.213, 5.300, 0.638, 4.305, 11.18};

void normalize3(std::array<double, 3>& v)
{
    double acc = 0.0;
    for(double c : v) acc = acc + c * c;             // vecMagAdd: x + y*y (UTIL/utilityFxns.hpp:21-23)
    const double norm = std::sqrt(acc);
    for(double& c : v) c = c / norm;
}

} // namespace

// ---------------------------------------------------------------------------------------------------
// Obj
// ---------------------------------------------------------------------------------------------------
Obj::Obj(SHAPE s, double eps, double mu, std::vector<LorenzDipoleOscillator> pols, bool ML, std::vector<double> geo,
         std::array<double, 3> loc, std::array<std::array<double, 3>, 3> uvec)
    : shape_(s), ML_(ML), eps_infty_(eps), mu_infty_(mu), geoParam_(geo), geoParamML_(geo), location_(loc), unitVec_(uvec), pols_(std::move(pols))
{
    // OBJECTS/Obj.cpp:74-83: coordTransform_[ii*3+jj] = (u_ii . e_jj) / |u_ii|   (NaN rows for the null z axis of 2-D runs, kept)
    for(int ii = 0; ii < 3; ++ii)
    {
        double mag2 = 0.0;
        for(double c : uvec[ii]) mag2 = mag2 + c * c;
        for(int jj = 0; jj < 3; ++jj)
        {
            double dot = 0.0;
            for(int k = 0; k < 3; ++k) dot = dot + uvec[ii][k] * (k == jj ? 1.0 : 0.0);
            coordTransform_[ii * 3 + jj] = dot / std::sqrt(mag2);
        }
    }
}

void Obj::setUpConsts(double dt)
{
    constsSet_ = true;
    for(const auto& pol : pols_)
    {
        if(pol.dipOrE_ != DIPOR::ISOTROPIC || pol.dipOrM_ != DIPOR::ISOTROPIC) useOrientedDipols_ = true;
        if(std::abs(pol.sigP_) != 0.0)
        {
            dipOr_.push_back(pol.dipOrE_);
            dipE_.push_back(pol.uVecDipE_);
            dipNormCompE_.push_back(pol.normCompWeightE_);
            dipTanLatCompE_.push_back(pol.tangentLatCompWeightE_);
            dipTanLongCompE_.push_back(pol.tangentLongCompWeightE_);
            alpha_.push_back(((2 - std::pow(pol.omg_ * dt, 2.0)) / (1 + pol.gam_ * dt)));
            xi_.push_back(((pol.gam_ * dt - 1) / (1 + pol.gam_ * dt)));
            gamma_.push_back(((pol.sigP_ * std::pow(pol.omg_ * dt, 2.0)) / (1 + pol.gam_ * dt)));
        }
        if(std::abs(pol.sigM_) != 0.0)
        {
            magAlpha_.push_back(((2 - std::pow(pol.omg_ * dt, 2.0)) / (1 + pol.gam_ * dt)));
            magXi_.push_back(((pol.gam_ * dt - 1) / (1 + pol.gam_ * dt)));
            magGamma_.push_back(((pol.sigM_ * std::pow(pol.omg_ * dt, 2.0)) / (1 + pol.gam_ * dt)));
        }
        if(std::abs(pol.tau_) != 0.0)
        {
            // the 1 / dt of the two gammas is the time derivative of the chiral interaction (OBJECTS/Obj.cpp:345-353)
            chiAlpha_.push_back(((2 - std::pow(pol.omg_ * dt, 2.0)) / (1 + pol.gam_ * dt)));
            chiXi_.push_back(((pol.gam_ * dt - 1) / (1 + pol.gam_ * dt)));
            chiGamma_.push_back((-1.0 / dt) * ((pol.tau_ * std::pow(pol.omg_ * dt, 2.0)) / ((1 + pol.gam_ * dt))));
            chiGammaPrev_.push_back((1.0 / dt) * ((pol.tau_ * std::pow(pol.omg_ * dt, 2.0)) / ((1 + pol.gam_ * dt))));
        }
    }
}

bool Obj::identityAxes() const
{
    for(int i = 0; i < 3; ++i)
        for(int j = 0; j < 3; ++j)
            if(coordTransform_[i * 3 + j] != (i == j ? 1.0 : 0.0)) return false;
    return true;
}

// Obj::findGradient of the shapes built here (OBJECTS/Obj.cpp:585-596 sphere, :644-653 cylinder, :655-666 block), for objects whose axes are the
// Cartesian ones: RealSpace2ObjectSpace is then the shift to the centre and ObjectSpace2RealSpace (the LAPACK inverse of the unit matrix) the
// normalisation alone.  (Rotated objects would need the reference's dgetrf / dgetri inverse bit for bit; the set-up refuses them.)
std::array<double, 3> Obj::findGradient(const std::array<double, 3>& pt) const
{
    auto magSq = [](const std::array<double, 3>& v) { double a = 0.0; for(double x : v) a = a + x * x; return a; };
    auto normalised = [&](std::array<double, 3> g) {
        if(magSq(g) < 1e-20) return std::array<double, 3>{{0.0, 0.0, 0.0}};
        const double norm = std::sqrt(magSq(g));
        for(double& x : g) x = x / norm;
        return g;
    };
    std::array<double, 3> v;
    for(int k = 0; k < 3; ++k) v[k] = pt[k] - location_[k];
    if(shape_ == SHAPE::SPHERE) return normalised(v);
    // (v_trans = 1 * v_cen + 0 + 0 in the reference's dgemv: exact)
    std::array<double, 3> grad = {{0.0, 0.0, 0.0}};
    if(shape_ == SHAPE::BLOCK)
    {
        std::array<double, 3> ptRat;
        for(int k = 0; k < 3; ++k) ptRat[k] = v[k] / geoParam_[k];
        int mx = 0;                                       // idamax_: first index of the largest magnitude
        for(int k = 1; k < 3; ++k) if(std::abs(ptRat[k]) > std::abs(ptRat[mx])) mx = k;
        for(int k = 0; k < 3; ++k)
            if(ptRat[k] == ptRat[mx]) grad[k] = v[k] >= 0 ? 1.0 : -1.0;
        return normalised(grad);
    }
    // cylinder
    const double r = std::sqrt(std::pow(v[0], 2.0) + std::pow(v[2], 2.0));
    const double len = geoParam_[1] * (r / geoParam_[0]);
    grad = {{v[0], 0.0, v[2]}};
    if(std::abs(v[1]) >= len / 2.0) grad = {{0.0, v[1] / std::abs(v[1]), 0.0}};
    return normalised(grad);
}

void Obj::addMLBuff(double d)
{
    // OBJECTS/Obj.hpp:537,653,949: sphere radius, block edges, cylinder radius + length grow by 3 d
    if(shape_ == SHAPE::SPHERE) geoParamML_[0] += 3.0 * d;
    else if(shape_ == SHAPE::BLOCK) { geoParamML_[0] += 3.0 * d; geoParamML_[1] += 3.0 * d; geoParamML_[2] += 3.0 * d; }
    else { geoParamML_[0] += 3.0 * d; geoParamML_[1] += 3.0 * d; }
}

bool Obj::isObj(const std::array<double, 3>& v, double dx, const std::vector<double>& geo) const
{
    if(shape_ == SHAPE::SPHERE)
    {
        // sphere::isObj, OBJECTS/Obj.cpp:373-380; dist :577-583
        double sum = 0;
        for(int cc = 0; cc < 3; ++cc) sum += std::pow((v[cc] - location_[cc]), 2);
        return !(std::sqrt(sum) > geo[0] + dx / 1.0e6);
    }
    // RealSpace2ObjectSpace (OBJECTS/Obj.cpp:13-22): v_trans = coordTransform^T-style projection on the object axes
    std::array<double, 3> v_cen, v_trans;
    for(int k = 0; k < 3; ++k) v_cen[k] = v[k] - location_[k];
    for(int j = 0; j < 3; ++j)
    {
        double temp = 0.0;
        for(int i = 0; i < 3; ++i) temp = temp + coordTransform_[i + j * 3] * v_cen[i];
        v_trans[j] = temp;
    }
    if(shape_ == SHAPE::BLOCK)
    {
        // block::isObj, OBJECTS/Obj.cpp:400-410 (comparisons written so that NaN coordinates pass, as there)
        for(int ii = 0; ii < 3; ++ii)
            if((v_trans[ii] > geo[ii] / 2.0 + dx / 1.0e6) || (v_trans[ii] < -1.0 * geo[ii] / 2.0 - dx / 1.0e6)) return false;
        return true;
    }
    // cylinder::isObj, OBJECTS/Obj.cpp:523-533
    if((v_trans[1] < -1.0 * geo[1] / 2.0 - dx * 1e-6) || (v_trans[1] > geo[1] / 2.0 + dx * 1e-6)) return false;
    double sum = 0;
    sum += std::pow(v_trans[0], 2); sum += std::pow(0.0, 2); sum += std::pow(v_trans[2], 2);
    if(std::sqrt(sum) > geo[0] + 1.0e-6 * dx) return false;
    return true;
}

std::array<double, 3> Obj::halfExtent(const std::vector<double>& geo) const
{
    const double inf = std::numeric_limits<double>::infinity();
    std::array<double, 3> h = {{inf, inf, inf}};
    if(shape_ == SHAPE::SPHERE) { h = {{geo[0], geo[0], geo[0]}}; return h; }
    for(double c : coordTransform_) if(!std::isfinite(c)) return h;   // degenerate axes (2-D): no culling along anything
    if(shape_ == SHAPE::BLOCK)
    {
        for(int a = 0; a < 3; ++a)
        {
            double e = 0.0;
            for(int i = 0; i < 3; ++i) e += std::abs(coordTransform_[i * 3 + a]) * geo[i] / 2.0;   // |u_i . e_a| * half edge
            h[a] = e;
        }
        return h;
    }
    const double rad = std::sqrt(geo[0] * geo[0] + geo[1] * geo[1] / 4.0);   // cylinder: bounding sphere
    h = {{rad, rad, rad}};
    return h;
}

// ---------------------------------------------------------------------------------------------------
// materials
// ---------------------------------------------------------------------------------------------------
std::vector<LorenzDipoleOscillator> Inputs::getMetal(const std::vector<double>& params) const
{
    // parallelInputs.cpp:1187-1221
    std::vector<LorenzDipoleOscillator> out;
    const double wp = ev2FDTD(params[0]);
    for(size_t ii = 0; ii < (params.size() - 1) / 3; ii++)
    {
        LorenzDipoleOscillator osc;
        const double f = params[3 * ii + 1];
        const double GAM = ev2FDTD(params[3 * ii + 2]);
        double OMG = ev2FDTD(params[3 * ii + 3]);
        if(OMG == 0.0) OMG = ev2FDTD(1.0e-20);
        osc.sigP_ = f * std::pow(wp / OMG, 2);
        osc.gam_ = GAM * M_PI;
        osc.omg_ = OMG * 2.0 * M_PI;
        out.push_back(osc);
    }
    return out;
}

bool Inputs::getMater(const std::string& mat, double& eps, double& mu, std::vector<LorenzDipoleOscillator>& pols) const
{
    // parallelInputs.cpp:1066-1180 (the entries the configurations use)
    mu = 1.0;
    pols.clear();
    if(mat == "Au" || mat == "au" || mat == "AU") { eps = 1.0 + 1.0e-14; pols = getMetal(AU_MAT); return true; }
    if(mat == "Ag" || mat == "ag" || mat == "AG") { eps = 1.0 + 1.1e-14; pols = getMetal(AG_MAT); return true; }
    if(mat == "Al" || mat == "al" || mat == "AL") { eps = 1.0 + 1.2e-14; pols = getMetal(AL_MAT); return true; }
    if(mat == "Cu" || mat == "cu" || mat == "CU") { eps = 1.0 + 1.3e-14; pols = getMetal(CU_MAT); return true; }
    if(mat == "TiO2") { eps = 6.20001; return true; }
    if(mat == "SiO2") { eps = 2.1025; return true; }
    if(mat == "Si") { eps = 11.8336; return true; }
    if(mat == "CdSe") { eps = 6.20; return true; }
    if(mat == "PbS") { eps = 17.20; return true; }
    if(mat == "vac" || mat == "Vac" || mat == "VAC") { eps = 1.0; return true; }
    return false;
}

std::shared_ptr<Obj> Inputs::jsonToObject(const Json& o)
{
    // parallelInputs.cpp:1263-1623
    const std::array<double, 3> loc = as_ptArr<double>(o, "loc");
    const std::string material = o.get<std::string>("material");
    double eps_infty = 0.0, mu_infty = 0.0;
    std::vector<LorenzDipoleOscillator> lorPols;
    if(material == "custom")
    {
        eps_infty = o.get<double>("eps", 1.00);
        mu_infty = o.get<double>("mu", 1.0);
        if(o.get<double>("tellegen", 0.0) != 0.0) throw std::logic_error("Tellegen media are outside the covered hot path");
        for(const auto& it : o.child("pols").kids)
        {
            const Json& p = it.second;
            LorenzDipoleOscillator osc;
            if(p.get<bool>("molecular_trans", false)) throw std::logic_error("molecular_trans poles are outside the covered hot path");
            const bool useTanIso = p.get<bool>("tanIso", false);
            osc.dipOrE_ = string2dipor(p.get<std::string>("dipOrE", "isotropic"));
            if(osc.dipOrE_ == DIPOR::LAT_TAN || osc.dipOrE_ == DIPOR::LONG_TAN)
                throw std::logic_error("lat_tangent / long_tangent dipole orientations are outside the covered hot path");
            if(useTanIso && (osc.dipOrE_ == DIPOR::ISOTROPIC || osc.dipOrE_ == DIPOR::UNIDIRECTIONAL))
                throw std::logic_error("Object can't be isotropic in tangential directions and unidirectional or isotropic");
            osc.gam_ = p.get<double>("gamma") * M_PI;
            osc.omg_ = p.get<double>("omega") * 2 * M_PI;
            osc.sigP_ = p.get<double>("sigma_p", 0.0);
            osc.sigM_ = p.get<double>("sigma_m", 0.0);
            osc.tau_ = p.get<double>("tau", 0.0);
            osc.dipOrM_ = string2dipor(p.get<std::string>("dipOrM", "isotropic"));
            if(useTanIso && (osc.dipOrM_ == DIPOR::ISOTROPIC || osc.dipOrM_ == DIPOR::UNIDIRECTIONAL))
                throw std::logic_error("Object can't be isotropic in tangential directions and unidirectional or isotropic");
            if((osc.sigM_ != 0.0 || osc.tau_ != 0.0) && (osc.dipOrE_ != DIPOR::ISOTROPIC || osc.dipOrM_ != DIPOR::ISOTROPIC))
                throw std::logic_error("oriented magnetic / chiral dipoles are outside the covered hot path");
            // (a pole without magnetic / chiral strength carries its dipOrM along unused: only useOrientedDipols_ sees it)
            // weights of the normal and the tangents (parallelInputs.cpp:1329-1357)
            {
                double polAngRelE = p.get<double>("polAngRelNormE", 45.0), azAngRelE = p.get<double>("azAngRelNormE", 45.0);
                const std::string how = p.get<std::string>("dipOrE", "isotropic");
                if(how == "normal") polAngRelE = 0.0;
                else if(how == "tangent") polAngRelE = 90.0;
                if(useTanIso && ((static_cast<int>(azAngRelE) % 45) != 0 || (static_cast<int>(azAngRelE) % 90) == 0))
                    throw std::logic_error("isotropic tangent is true, but azimuthal angle of electric dipole is not 45 degrees");
                osc.normCompWeightE_ = std::cos(polAngRelE * M_PI / 180.0);
                osc.tangentLatCompWeightE_ = std::sin(polAngRelE * M_PI / 180.0) * std::sin(azAngRelE * M_PI / 180.0);
                osc.tangentLongCompWeightE_ = std::sin(polAngRelE * M_PI / 180.0) * std::cos(azAngRelE * M_PI / 180.0);
            }
            if(osc.dipOrE_ == DIPOR::UNIDIRECTIONAL)
            {
                if(osc.sigP_ > 0.0) { osc.uVecDipE_ = as_ptArr<double>(p, "dirDipE"); normalize3(osc.uVecDipE_); }
                else osc.uVecDipE_ = {{0.0, 0.0, 0.0}};
            }
            else osc.uVecDipE_ = {{1.0, 1.0, 1.0}};
            if(useTanIso)
            {
                // one pole along each tangent, the normal weight shared (parallelInputs.cpp:1412-1428)
                LorenzDipoleOscillator oscTanLong(osc), oscTanLat(osc);
                oscTanLong.tangentLatCompWeightE_ = 0.0;
                oscTanLat.tangentLongCompWeightE_ = 0.0;
                oscTanLong.normCompWeightE_ /= std::sqrt(2.0);
                oscTanLat.normCompWeightE_ /= std::sqrt(2.0);
                lorPols.push_back(oscTanLat);
                lorPols.push_back(oscTanLong);
            }
            else
            lorPols.push_back(osc);
        }
    }
    else if(!getMater(material, eps_infty, mu_infty, lorPols))
        throw std::logic_error("The material name " + material + " is not available in this build");

    std::array<std::array<double, 3>, 3> unitVecs;
    const Json& uvecs = o.child("unit_vectors");
    int cc = 0;
    for(const auto& it : uvecs.kids) { if(cc < 3) unitVecs[cc] = as_ptArr<double>(it.second, "uvec"); ++cc; }
    if(uvecs.size() == 0)
    {
        const double orTheta = M_PI / 2.0 - o.get<double>("orTheta", 90.0) * M_PI / 180.0;
        const double orPhi = -1.0 * o.get<double>("orPhi", 0.0) * M_PI / 180.0;
        if(size_[2] == 0)
        {
            unitVecs[0] = {{std::cos(orPhi), -1.0 * std::sin(orPhi), 0}};
            unitVecs[1] = {{std::sin(orPhi), std::cos(orPhi), 0}};
            unitVecs[2] = {{0.0, 0.0, 0}};
        }
        else
        {
            unitVecs[0] = {{std::cos(orTheta) * std::cos(orPhi), -1.0 * std::cos(orTheta) * std::sin(orPhi), std::sin(orTheta)}};
            unitVecs[1] = {{std::sin(orPhi), std::cos(orPhi), 0}};
            unitVecs[2] = {{-1.0 * std::sin(orTheta) * std::cos(orPhi), std::sin(orTheta) * std::sin(orPhi), std::cos(orTheta)}};
        }
    }
    const bool ML = o.child("Basis_Set").size() > 0;
    const std::string shape = o.get<std::string>("shape");
    if(shape == "block")
    {
        std::vector<double> geo = as_vector<double>(o, "size");
        geo.push_back(o.get<double>("rad_curve", 0.0));
        if(geo.back() != 0.0) throw std::logic_error("rounded blocks are outside the covered hot path");
        return std::make_shared<Obj>(SHAPE::BLOCK, eps_infty, mu_infty, lorPols, ML, geo, loc, unitVecs);
    }
    if(shape == "sphere")
        return std::make_shared<Obj>(SHAPE::SPHERE, eps_infty, mu_infty, lorPols, ML, std::vector<double>{o.get<double>("radius", 0.0)}, loc, unitVecs);
    if(shape == "cylinder")
        return std::make_shared<Obj>(SHAPE::CYLINDER, eps_infty, mu_infty, lorPols, ML,
                                     std::vector<double>{o.get<double>("radius", 0.0), o.get<double>("length", 0.0)}, loc, unitVecs);
    throw std::logic_error("shape " + shape + " is outside the covered hot path (block, sphere, cylinder)");
}

// ---------------------------------------------------------------------------------------------------
// Inputs
// ---------------------------------------------------------------------------------------------------
Inputs::Inputs(const Json& IP, bool postProcessing)
{
    periodic_ = IP.get<bool>("CompCell.PBC", false);
    pol_ = string2pol(IP.get<std::string>("CompCell.pol"));
    res_ = IP.get<int>("CompCell.res", -1);
    courant_ = IP.get<double>("CompCell.courant", 0.5);
    a_ = IP.get<double>("CompCell.a", 1e-7);
    tMax_ = IP.get<double>("CompCell.tLim");
    I0_ = IP.get<double>("CompCell.I0", a_ * EPS0() * SPEED_OF_LIGHT);
    // periodic boundaries with a k-point switch the reference to complex fields (INPUTS/parallelInputs.cpp:108-112): real fields only here
    // periodic boundaries with a k-point switch the reference to complex fields (INPUTS/parallelInputs.cpp:22-25,108-112)
    k_point_ = as_ptArr<double>(IP, "CompCell.k-point", 0.0);
    cplxFields_ = IP.get<bool>("CompCell.cplxFields", false);
    if(periodic_)
        for(double kk : k_point_) if(kk != 0) cplxFields_ = true;
    if(cplxFields_ && !periodic_)
        throw std::logic_error("complex fields without periodic boundaries are outside the covered hot path (SURVEY.md section 8(f) rank 3)");
    size_ = as_ptArr<double>(IP, "CompCell.size");
    d_ = as_ptArr<double>(IP, "CompCell.stepSize", 1.0 / res_);
    dt_ = IP.get<double>("CompCell.dt", courant_ / std::sqrt(1.0 / (d_[0] * d_[0]) + 1.0 / (d_[1] * d_[1]) + 1.0 / (d_[2] * d_[2])));
    pmlSigOptRat_ = IP.get<double>("PML.sigOptRat", 1.0);
    pmlKappaMax_ = IP.get<double>("PML.kappaMax", 1.0);
    pmlAMax_ = IP.get<double>("PML.aMax", 0.00);
    pmlMa_ = IP.get<double>("PML.ma", 1.0);
    pmlM_ = IP.get<double>("PML.m", 3.0);
    if(d_[0] < 0 || d_[1] < 0 || d_[2] < 0) throw std::logic_error("Please define a positive step size.");
    if(dt_ > 1.0 / std::sqrt(std::accumulate(d_.begin(), d_.end(), 0.0, [](double a, double b) { return a + 1.0 / (b * b); })))
        throw std::logic_error("Time step is larger than the stable time step for the calculation.");
    const std::array<double, 3> pmlThickness = as_ptArr<double>(IP, "PML.thickness");
    for(int ii = 0; ii < 3; ++ii)
    {
        if(pmlThickness[ii] * 2 > size_[ii]) throw std::logic_error("PML size is larger than the cell size, this will lead to infinte fields.");
        pmlThickness_[ii] = find_pt(pmlThickness[ii], d_[ii]);
    }
    if(IP.child("TFSF").size() > 0 && !postProcessing) throw std::logic_error("TFSF sources are outside the covered hot path (SURVEY.md section 8(f) rank 1)");

    // ---- sources (parallelInputs.cpp:114-216) ----
    for(const auto& it : IP.child("SourceList").kids)
    {
        const Json& s = it.second;
        SourceInput src;
        src.pol = string2pol(s.get<std::string>("pol"));
        if(src.pol == POLARIZATION::L || src.pol == POLARIZATION::R) throw std::logic_error("circularly polarised sources are outside the covered hot path");
        for(const auto& pit : s.child("PulseList").kids)
        {
            const Json& pul = pit.second;
            const PLSSHAPE shape = string2prof(pul.get<std::string>("profile"));
            const double emax = pul.get<double>("Field_Intensity", 1.0) * a_ * EPS0() * SPEED_OF_LIGHT / I0_;
            std::vector<cplx> fxn;
            switch(shape)
            {
                case PLSSHAPE::GAUSSIAN:
                    fxn.push_back(cplx(0.0, -1.0 * pul.get<double>("fcen") * 2 * M_PI));
                    fxn.push_back(1.0 / pul.get<double>("fwidth"));
                    fxn.push_back(pul.get<double>("cutoff") * fxn[1]);
                    fxn.push_back(pul.get<double>("t_0", std::real(fxn[1] * fxn[2])));
                    fxn[2] += fxn[3];
                    break;
                case PLSSHAPE::BH:
                    fxn.push_back(cplx(0.0, -1.0 * pul.get<double>("fcen") * 2 * M_PI));
                    fxn.push_back(pul.get<double>("tau", 1.0 / pul.get<double>("fwidth")));
                    fxn.push_back(pul.get<double>("t_0", std::real(fxn[1] * fxn[2])));
                    fxn.push_back(pul.get<double>("BH1", 0.35875));
                    fxn.push_back(pul.get<double>("BH2", 0.48829));
                    fxn.push_back(pul.get<double>("BH3", 0.14128));
                    fxn.push_back(pul.get<double>("BH4", 0.01168));
                    break;
                case PLSSHAPE::RECT:
                    fxn.push_back(pul.get<double>("tau"));
                    fxn.push_back(pul.get<double>("t_0"));
                    fxn.push_back(pul.get<double>("n", 30) * 2.0);
                    fxn.push_back(cplx(0.0, -1.0 * pul.get<double>("fcen") * 2 * M_PI));
                    break;
                case PLSSHAPE::CONTINUOUS:
                    fxn.push_back(cplx(0.0, -1.0 * pul.get<double>("fcen") * 2 * M_PI));
                    break;
                case PLSSHAPE::RAMP_CONT:
                    fxn.push_back(cplx(0.0, -1.0 * pul.get<double>("fcen") * 2 * M_PI));
                    fxn.push_back(pul.get<double>("ramp_val"));
                    break;
                case PLSSHAPE::RICKER:
                    fxn.push_back(cplx(0.0, -1.0 * pul.get<double>("fcen") * 2 * M_PI));
                    fxn.push_back(pul.get<double>("fwidth"));
                    fxn.push_back(pul.get<double>("cutoff"));
                    break;
            }
            fxn.push_back(emax);   // parallelFDTDField.cpp:476
            src.shapes.push_back(shape);
            src.fxn.push_back(fxn);
        }
        const std::array<double, 3> tempSz = as_ptArr<double>(s, "size");
        for(int c = 0; c < 3; ++c) src.sz[c] = find_pt(tempSz[c], d_[c]) + 1;
        src.phi = s.get<double>("phi", 90);
        if(!(src.phi == 90 || src.phi == 180 || src.phi == 270 || src.phi == 0))
            throw std::logic_error("oblique sources are outside the covered hot path");
        const std::array<double, 3> locs = as_ptArr<double>(s, "loc");
        for(int i = 0; i < 3; ++i)
        {
            if(locs[i] + tempSz[i] / 2.0 > size_[i] / 2.0 || locs[i] - tempSz[i] / 2.0 < -1.0 * size_[i] / 2.0)
                throw std::logic_error("The source is at least partially outside the FDTD Cell");
            src.loc[i] = find_pt(locs[i] + size_[i] / 2.0 - tempSz[i] / 2.0, d_[i]);
        }
        sources_.push_back(src);
    }

    // ---- objects (parallelInputs.cpp:410-666) ----
    std::array<std::array<double, 3>, 3> uVecs;
    for(int ii = 0; ii < 3; ++ii) { uVecs[ii] = {{0.0, 0.0, 0.0}}; uVecs[ii][ii] = 1.0; }
    objArr_.push_back(std::make_shared<Obj>(SHAPE::BLOCK, 1.0, 1.0, std::vector<LorenzDipoleOscillator>(), false,
                                            std::vector<double>{size_[0], size_[1], size_[2]}, std::array<double, 3>{{0.0, 0.0, 0.0}}, uVecs));
    int qq = 0;
    for(const auto& it : IP.child("ObjectList").kids)
    {
        const Json& o = it.second;
        std::shared_ptr<Obj> obj = jsonToObject(o);
        std::vector<std::array<int, 2>> basis;
        for(const auto& b : o.child("Basis_Set").kids) basis.push_back({{b.second.get<int>("l"), b.second.get<int>("m")}});
        const int nx = find_pt(size_[0], d_[0]) + 1, ny = find_pt(size_[1], d_[1]) + 1, nz = find_pt(size_[2], d_[2]) + 1;
        const double dmin = *std::min_element(d_.begin(), d_.end());
        // later objects take emitter nodes away from earlier emitter objects (:439-451)
        for(auto& q : qes_)
            for(size_t ll = 0; ll < q.locs.size(); ll++)
            {
                std::array<double, 3> loc = {{static_cast<double>(q.locs[ll][0] - (nx - 1) / 2.0 - 1) * d_[0], static_cast<double>(q.locs[ll][1] - (ny - 1) / 2.0 - 1) * d_[1],
                                              static_cast<double>(q.locs[ll][2] - (nz - 1) / 2.0 - 1) * d_[2]}};
                if(obj->isObj(loc, dmin, obj->geoParam_)) { q.locs.erase(q.locs.begin() + ll); --ll; }
            }
        const double e_conv = ELEMENTARY_CHARGE * EPS0() * std::pow(SPEED_OF_LIGHT / I0_, 2) / a_;
        if(basis.size() > 0)
        {
            QEInput q;
            q.object = (int)objArr_.size();
            q.basis = basis;
            q.density = o.get<double>("mol_den", 1.0);
            for(const auto& eit : o.child("Energy_Levels").kids)
            {
                const Json& eLev = eit.second;
                EnergyLevel level;
                std::vector<double> e_cen = as_vector<double>(eLev, "E_cen");
                for(double& e : e_cen) e = e_conv * e;
                std::vector<double> weight_vec = as_vector<double>(eLev, "weights", 1.0, (int)e_cen.size());
                const std::string dist = eLev.get<std::string>("distribution", "Delta_Fxn");
                level.nstates_ = eLev.get<int>("nstates", 1);
                const bool delta = dist == "delta_fxn" || dist == "Delta_fxn" || dist == "Delta_Fxn" || dist == "delta_Fxn";
                if(!(delta || level.nstates_ == 1)) throw std::logic_error("broadened energy-level distributions are outside the covered hot path");
                if(level.nstates_ > 1) throw std::logic_error("The number of states is greater than one for a delta function distribution.");
                level.weights_ = weight_vec;
                level.energyStates_ = e_cen;
                level.levDescribed_ = eLev.get<int>("levs_described", 1);
                q.levels.push_back(level);
            }
            for(const auto& c : o.child("couplings").kids) q.couplings.push_back(c.second.value<double>() * 1.0e-21 / std::pow(a_, 2.0) / I0_);
            for(const auto& gv : o.child("gam").kids)
            {
                std::vector<double> gam;
                for(const auto& g : gv.second.child("g").kids) gam.push_back(g.second.value<double>() * a_ / SPEED_OF_LIGHT);
                q.gam.push_back(gam);
            }
            // same scan order as the reference (z outermost, x innermost, :596-614), restricted to the object's bounding box
            int lo[3] = {0, 0, 0}, hi[3] = {nx, ny, nz};
            {
                const std::array<double, 3> h = obj->halfExtent(obj->geoParam_);
                const int nn[3] = {nx, ny, nz};
                for(int k = 0; k < 3; ++k)
                    if(std::isfinite(h[k]))
                    {
                        const double c = (nn[k] - 1) / 2.0;
                        lo[k] = std::max(0, (int)std::floor((obj->location_[k] - h[k]) / d_[k] + c) - 2);
                        hi[k] = std::min(nn[k], (int)std::ceil((obj->location_[k] + h[k]) / d_[k] + c) + 3);
                    }
            }
            for(int zz = lo[2]; zz < hi[2]; ++zz)
                for(int yy = lo[1]; yy < hi[1]; ++yy)
                    for(int xx = lo[0]; xx < hi[0]; ++xx)
                    {
                        std::array<double, 3> loc = {{static_cast<double>(xx - (nx - 1) / 2.0) * d_[0], static_cast<double>(yy - (ny - 1) / 2.0) * d_[1],
                                                      static_cast<double>(zz - (nz - 1) / 2.0) * d_[2]}};
                        if(obj->isObj(loc, dmin, obj->geoParam_)) q.locs.push_back({{xx, yy, zz}});
                    }
            for(const auto& r : o.child("RelaxationOperators").kids)
                q.relax.push_back({r.second.get<int>("state_i"), r.second.get<int>("state_f"), r.second.get<double>("rate") * a_ / SPEED_OF_LIGHT,
                                   r.second.get<double>("dephasing_rate", 0.0) * a_ / SPEED_OF_LIGHT});
            if(o.get<bool>("output_pol", false)) throw std::logic_error("emitter polarisation output files are outside the covered hot path");
            for(const auto& l : o.child("dtc_levs").kids) q.dtcLevs.push_back(l.second.value<int>());
            q.dtcPopTimeInt = o.get<int>("levDTC_timeInt", 1);
            for(int lev : q.dtcLevs)
                q.dtcPopFiles.push_back(o.get<std::string>("dtc_pop_fname_base", "output_data/qe_") + std::to_string(qq) + "_level_" + std::to_string(lev) + ".dat");
            obj->addMLBuff(*std::max_element(d_.begin(), d_.end()));
            qes_.push_back(q);
            ++qq;
        }
        objArr_.push_back(obj);
    }

    // ---- detectors (parallelInputs.cpp:671-755) ----
    int ii = 0;
    for(const auto& it : IP.child("DetectorList").kids)
    {
        const Json& dj = it.second;
        DetectorInput d;
        d.type = string2out(dj.get<std::string>("type"));
        d.name = dj.get<std::string>("fname") + "_field_" + std::to_string(ii) + ".dat";
        ++ii;
        d.cls = string2dtcclass(dj.get<std::string>("dtc_class", "cout"));
        if(d.cls == DTCCLASS::BMP) throw std::logic_error("bmp detectors are outside the covered hot path");
        d.SI = dj.get<bool>("SI", true);
        d.timeInt = dj.get<double>("Time_Interval", dt_);
        const std::array<double, 3> tempSz = as_ptArr<double>(dj, "size");
        for(int c = 0; c < 3; ++c) d.sz[c] = find_pt(tempSz[c], d_[c]) + 1;
        const std::array<double, 3> locs = as_ptArr<double>(dj, "loc");
        for(int i = 0; i < 3; ++i)
        {
            if(locs[i] - tempSz[i] / 2.0 < -1.0 * size_[i] / 2.0 || locs[i] + tempSz[i] / 2.0 > size_[i] / 2.0) throw std::logic_error("A detector is outside the FDTD cell.");
            d.loc[i] = find_pt(locs[i] + size_[i] / 2.0 - tempSz[i] / 2.0, d_[i]);
        }
        if(d.cls == DTCCLASS::FREQ)
        {
            // the frequency list (parallelInputs.cpp:716-751)
            FreqDtcInput q;
            q.type = d.type; q.SI = d.SI; q.name = d.name; q.loc = d.loc; q.sz = d.sz;
            q.outputMaps = dj.get<bool>("output_map", false);
            q.timeInt = static_cast<int>(std::floor(d.timeInt / dt_ + 0.5));
            const double fCen = dj.get<double>("fcen", -1.0), fWidth = dj.get<double>("fwidth", -1.0);
            const double lamL = dj.get<double>("lamL", -1.0), lamR = dj.get<double>("lamR", -1.0);
            const int nFreq = dj.get<int>("nfreq", -1);
            if(nFreq < 1) throw std::logic_error("The freq detector regions need to have a number of frequencies specified");
            q.freqs.assign(nFreq, 0.0);
            if(fCen != -1.0 && fWidth != -1.0)
            {
                if(lamL != -1.0 && lamR != -1.0) throw std::logic_error("Both a freq and wavelength range is defined, please select one to define for the freq detector");
                const double dOmg = fWidth / static_cast<double>(nFreq - 1);
                for(int k = 0; k < nFreq; ++k) q.freqs[k] = (fCen - fWidth / 2.0 + k * dOmg) * 2.0 * M_PI;
            }
            else if(lamL != -1.0 && lamR != -1.0)
            {
                const double dLam = (lamR - lamL) / static_cast<double>(nFreq - 1);
                for(int k = 0; k < nFreq; ++k) q.freqs[k] = 2.0 * M_PI / (lamL + k * dLam);
            }
            else throw std::logic_error("All frequency detectors must either have fcen and fwidth defined or lamL and lamR defined");
            if(q.timeInt < 1) throw std::logic_error("The time step of a detector is less than the main grid or set to 0.");
            freqDtcs_.push_back(q);
            continue;
        }
        detectors_.push_back(d);
    }
    // ---- flux regions (parallelInputs.cpp:757-838) ----
    for(const auto& it : IP.child("FluxList").kids)
    {
        const Json& fj = it.second;
        FluxInput f;
        f.name = fj.get<std::string>("name");
        const std::array<double, 3> tempSz = as_ptArr<double>(fj, "size");
        for(int c = 0; c < 3; ++c) f.sz[c] = find_pt(tempSz[c], d_[c]) + 1;
        const std::array<double, 3> locs = as_ptArr<double>(fj, "loc");
        for(int i = 0; i < 3; ++i) f.loc[i] = find_pt(locs[i] + size_[i] / 2.0 - tempSz[i] / 2, d_[i]);
        f.weight = fj.get<double>("weight", 1.0);
        f.timeInt = static_cast<int>(std::floor(fj.get<double>("Time_Interval", dt_) / (dt_) + 0.50));
        if(f.timeInt <= 0) f.timeInt = 1;
        const double fCen = fj.get<double>("fcen", -1.0), fWidth = fj.get<double>("fwidth", -1.0);
        const double lamL = fj.get<double>("lamL", -1.0), lamR = fj.get<double>("lamR", -1.0);
        const int nFreq = fj.get<int>("nfreq", -1);
        if(nFreq <= 0) throw std::logic_error("The flux regions need to have a number of frequencies specified");
        f.freqs.assign(nFreq, 0.0);
        if(fCen != -1.0 && fWidth != -1.0)
        {
            const double dOmg = fWidth / static_cast<double>(nFreq - 1);
            for(int k = 0; k < nFreq; ++k) f.freqs[k] = (fCen - fWidth / 2.0 + k * dOmg) * 2.0 * M_PI;
        }
        else if(lamL != -1.0 && lamR != -1.0)
        {
            const double dLam = (lamR - lamL) / static_cast<double>(nFreq - 1);
            for(int k = 0; k < nFreq; ++k) f.freqs[k] = 2.0 * M_PI / (lamL + k * dLam);
        }
        else throw std::logic_error("All fluxes must either have fcen and fwidth defined or lamL and lamR defined");
        f.SI = fj.get<bool>("SI", false);
        f.crossSec = fj.get<bool>("cross_sec", false);
        f.save = fj.get<bool>("save", false);
        f.load = fj.get<bool>("load", false);
        f.incdFile = fj.get<std::string>("incd_fileds", "");           // the reference's spelling (INPUTS/parallelInputs.cpp:828)
        if(f.load && f.incdFile.empty()) throw std::logic_error("Trying to load in file without a valid path, in the " + std::to_string(fluxes_.size()) + " flux detector");
        fluxes_.push_back(f);
    }
}

// UTIL/PulseFxn.hpp:19-90 with the same std::complex expressions
cplx pulseValue(PLSSHAPE shape, double tt, const std::vector<cplx>& param)
{
    switch(shape)
    {
        case PLSSHAPE::GAUSSIAN:
            return (tt <= std::real(param[2])) ? param[4] * exp(param[0] * tt - (pow((tt - param[3]) / param[1], 2.0) / 2.0)) : 0.0;
        case PLSSHAPE::CONTINUOUS:
            return param[1] * std::exp(param[0] * tt);
        case PLSSHAPE::RAMP_CONT:
            return (std::abs(param[2] * param[1] * tt) <= std::abs(param[2])) ? std::abs(param[1] * tt) * param[2] * std::exp(param[0] * tt) : param[2] * std::exp(param[0] * tt);
        case PLSSHAPE::BH:
            return (std::real(param[2] - param[1] / 2.0) <= tt && tt <= std::real(param[2] + param[1] / 2.0))
                       ? param[7] * std::exp(param[0] * (tt - param[2])) * (param[3] + param[4] * cos(2.0 * M_PI * (tt - param[2]) / param[1]) + param[5] * cos(4.0 * M_PI * (tt - param[2]) / param[1]) + param[6] * cos(6.0 * M_PI * (tt - param[2]) / param[1]))
                       : 0.0;
        case PLSSHAPE::RECT:
            return (tt < std::real(param[1] + param[0]) && tt > std::real(param[1] - param[0])) ? param[4] / (std::pow(2.0 * (tt - param[1]) / param[0], param[2]) + 1.0) * (std::exp(param[3] * (tt - param[1]))) : 0.0;
        case PLSSHAPE::RICKER:
            return (tt < std::real(param[2] * param[1] / param[0])) ? param[3] * (1.0 - 2.0 * std::pow(M_PI * (tt * param[0] - param[1]), 2.0)) * std::exp(-std::pow(M_PI * (tt * param[0] - param[1]), 2.0)) : 0.0;
    }
    return 0.0;
}

} // namespace chiml_host
