// rs_model_compile.cpp -- native MJCF -> rs_model_desc compiler behind the C ABI (rs_model_compile).
//
// Host-only C++ (no CUDA, no third-party XML library).  It is the native twin of roboschool_b200/mjcf.py with that
// module's DEFAULT conventions, and replaces what the reference obtains from its Bullet dependency through
//   World::load_mjcf -> b3LoadMJCFCommandInit(URDF_USE_SELF_COLLISION | ..._EXCLUDE_ALL_PARENTS)
//   (roboschool/cpp-household/physics-bullet.cpp:95-131) and b3GetJointInfo / b3GetBodyInfo (:206-251),
// so that a native binding can obtain a descriptor without Python.  The importer conventions are the
// ones listed in the header of mjcf.py (dummy links for multi-joint bodies, fixed massless base for jointed roots,
// COM at the last fromto capsule, AABB-box inertia, 1000 kg/m^3, OR collision filter, parent exclusion); the task
// constants (gym_mujoco_walkers.py class attributes) are NOT part of the MJCF and stay zero here: the caller fills them
// (roboschool_b200/envspec.py does it for the shipped envs).
//
// tests/test_model_compile_native.py checks this compiler field by field against mjcf.py on the shipped models and on
// synthetic MJCFs (boxes, rotated capsules, several jointed top-level bodies, radian models).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/roboschool_b200.h"

int rs_set_error(const char* s);   // rs_step.cu: the library's last-error slot

namespace {

// ------------------------------------------------------------------------------------------------------------------
// a small XML reader: elements, attributes, comments, processing instructions; text content is skipped
// ------------------------------------------------------------------------------------------------------------------
struct Xml {
    std::string tag;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<Xml> kids;
    const std::string* get(const char* name) const {
        for (const auto& a : attrs) if (a.first == name) return &a.second;
        return nullptr;
    }
    const Xml* child(const char* t) const {
        for (const auto& k : kids) if (k.tag == t) return &k;
        return nullptr;
    }
};

struct XmlReader {
    const std::string& s;
    size_t p = 0;
    explicit XmlReader(const std::string& src) : s(src) {}
    [[noreturn]] void fail(const char* what) const {
        throw std::runtime_error(std::string("XML: ") + what + " at byte " + std::to_string(p));
    }
    bool starts(const char* lit) const { return s.compare(p, strlen(lit), lit) == 0; }
    void skip_ws() { while (p < s.size() && isspace((unsigned char)s[p])) p++; }
    void skip_until(const char* lit) {
        size_t q = s.find(lit, p);
        if (q == std::string::npos) fail("unterminated construct");
        p = q + strlen(lit);
    }
    // skips text, comments, <? ?> and <! > up to the next element start or end tag; false at end of input
    bool next_markup() {
        for (;;) {
            while (p < s.size() && s[p] != '<') p++;
            if (p >= s.size()) return false;
            if (starts("<!--")) { skip_until("-->"); continue; }
            if (starts("<?")) { skip_until("?>"); continue; }
            if (starts("<!")) { skip_until(">"); continue; }
            return true;
        }
    }
    std::string name() {
        size_t b = p;
        while (p < s.size() && (isalnum((unsigned char)s[p]) || s[p] == '_' || s[p] == '-' || s[p] == ':' || s[p] == '.')) p++;
        if (p == b) fail("name expected");
        return s.substr(b, p - b);
    }
    Xml element() {     // p at '<' of a start tag
        Xml e;
        p++;
        e.tag = name();
        for (;;) {
            skip_ws();
            if (p >= s.size()) fail("unterminated tag");
            if (s[p] == '/') { if (!starts("/>")) fail("'/>' expected"); p += 2; return e; }
            if (s[p] == '>') { p++; break; }
            std::string k = name();
            skip_ws();
            if (p >= s.size() || s[p] != '=') fail("'=' expected");
            p++;
            skip_ws();
            if (p >= s.size() || (s[p] != '"' && s[p] != '\'')) fail("quoted value expected");
            const char q = s[p++];
            size_t b = p;
            while (p < s.size() && s[p] != q) p++;
            if (p >= s.size()) fail("unterminated attribute value");
            e.attrs.emplace_back(k, s.substr(b, p - b));
            p++;
        }
        for (;;) {
            if (!next_markup()) fail("missing end tag");
            if (starts("</")) {
                p += 2;
                std::string t = name();
                if (t != e.tag) fail("mismatched end tag");
                skip_ws();
                if (p >= s.size() || s[p] != '>') fail("'>' expected");
                p++;
                return e;
            }
            e.kids.push_back(element());
        }
    }
    Xml document() {
        if (!next_markup() || starts("</")) fail("no root element");
        return element();
    }
};

// ------------------------------------------------------------------------------------------------------------------
// small fixed-size linear algebra (double)
// ------------------------------------------------------------------------------------------------------------------
struct V3 { double v[3] = {0, 0, 0}; double& operator[](int i) { return v[i]; } double operator[](int i) const { return v[i]; } };
struct M3 { double m[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}; };
V3 operator+(const V3& a, const V3& b) { V3 r; for (int i = 0; i < 3; i++) r[i] = a[i] + b[i]; return r; }
V3 operator-(const V3& a, const V3& b) { V3 r; for (int i = 0; i < 3; i++) r[i] = a[i] - b[i]; return r; }
V3 operator*(double k, const V3& a) { V3 r; for (int i = 0; i < 3; i++) r[i] = k * a[i]; return r; }
V3 mul(const M3& R, const V3& a) { V3 r; for (int i = 0; i < 3; i++) r[i] = R.m[i][0] * a[0] + R.m[i][1] * a[1] + R.m[i][2] * a[2]; return r; }
V3 mul_abs(const M3& R, const V3& a) { V3 r; for (int i = 0; i < 3; i++) r[i] = std::fabs(R.m[i][0]) * a[0] + std::fabs(R.m[i][1]) * a[1] + std::fabs(R.m[i][2]) * a[2]; return r; }
M3 transpose(const M3& R) { M3 T; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T.m[i][j] = R.m[j][i]; return T; }
double norm(const V3& a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

std::vector<double> numbers(const std::string& s) {
    std::vector<double> out;
    const char* c = s.c_str();
    for (;;) {
        while (*c && isspace((unsigned char)*c)) c++;
        if (!*c) break;
        char* end = nullptr;
        double v = strtod(c, &end);
        if (end == c) throw std::runtime_error("not a number: '" + s + "'");
        out.push_back(v);
        c = end;
    }
    return out;
}
std::vector<double> numbers_n(const std::string& s, size_t n) {
    std::vector<double> v = numbers(s);
    if (v.size() != n) throw std::runtime_error("expected " + std::to_string(n) + " numbers, got '" + s + "'");
    return v;
}
V3 vec3_or(const std::string* s, double a, double b, double c) {
    V3 r; r[0] = a; r[1] = b; r[2] = c;
    if (s) { auto v = numbers_n(*s, 3); r[0] = v[0]; r[1] = v[1]; r[2] = v[2]; }
    return r;
}
M3 quat_wxyz_to_mat(const std::vector<double>& q0) {
    const double n = std::sqrt(q0[0] * q0[0] + q0[1] * q0[1] + q0[2] * q0[2] + q0[3] * q0[3]);
    const double w = q0[0] / n, x = q0[1] / n, y = q0[2] / n, z = q0[3] / n;
    M3 R;
    R.m[0][0] = 1 - 2 * (y * y + z * z); R.m[0][1] = 2 * (x * y - z * w);     R.m[0][2] = 2 * (x * z + y * w);
    R.m[1][0] = 2 * (x * y + z * w);     R.m[1][1] = 1 - 2 * (x * x + z * z); R.m[1][2] = 2 * (y * z - x * w);
    R.m[2][0] = 2 * (x * z - y * w);     R.m[2][1] = 2 * (y * z + x * w);     R.m[2][2] = 1 - 2 * (x * x + y * y);
    return R;
}
M3 axisangle_to_mat(const std::vector<double>& aa, bool degrees) {
    const double n = std::sqrt(aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2]);
    const double x = aa[0] / n, y = aa[1] / n, z = aa[2] / n;
    const double ang = degrees ? aa[3] * (M_PI / 180.0) : aa[3];
    const double c = std::cos(ang), s = std::sin(ang), C = 1 - c;
    M3 R;
    R.m[0][0] = c + x * x * C;     R.m[0][1] = x * y * C - z * s; R.m[0][2] = x * z * C + y * s;
    R.m[1][0] = y * x * C + z * s; R.m[1][1] = c + y * y * C;     R.m[1][2] = y * z * C - x * s;
    R.m[2][0] = z * x * C - y * s; R.m[2][1] = z * y * C + x * s; R.m[2][2] = c + z * z * C;
    return R;
}
// rotation matrix -> quaternion (x, y, z, w), w >= 0, normalised
void mat_to_quat_xyzw(const M3& Rm, double* q) {
    const double (*R)[3] = Rm.m;
    const double t = R[0][0] + R[1][1] + R[2][2];
    double w, x, y, z;
    if (t > 0) {
        const double s = std::sqrt(t + 1.0) * 2;
        w = 0.25 * s; x = (R[2][1] - R[1][2]) / s; y = (R[0][2] - R[2][0]) / s; z = (R[1][0] - R[0][1]) / s;
    } else if (R[0][0] > R[1][1] && R[0][0] > R[2][2]) {
        const double s = std::sqrt(1.0 + R[0][0] - R[1][1] - R[2][2]) * 2;
        w = (R[2][1] - R[1][2]) / s; x = 0.25 * s; y = (R[0][1] + R[1][0]) / s; z = (R[0][2] + R[2][0]) / s;
    } else if (R[1][1] > R[2][2]) {
        const double s = std::sqrt(1.0 + R[1][1] - R[0][0] - R[2][2]) * 2;
        w = (R[0][2] - R[2][0]) / s; x = (R[0][1] + R[1][0]) / s; y = 0.25 * s; z = (R[1][2] + R[2][1]) / s;
    } else {
        const double s = std::sqrt(1.0 + R[2][2] - R[0][0] - R[1][1]) * 2;
        w = (R[1][0] - R[0][1]) / s; x = (R[0][2] + R[2][0]) / s; y = (R[1][2] + R[2][1]) / s; z = 0.25 * s;
    }
    if (w < 0) { w = -w; x = -x; y = -y; z = -z; }
    const double n = std::sqrt(x * x + y * y + z * z + w * w);
    q[0] = x / n; q[1] = y / n; q[2] = z / n; q[3] = w / n;
}

// ------------------------------------------------------------------------------------------------------------------
// MJCF elements
// ------------------------------------------------------------------------------------------------------------------
constexpr double MARGIN = 0.04;            // CONVEX_DISTANCE_MARGIN (btCollisionMargin.h)
constexpr double BREAKING = 0.02;          // gContactBreakingThreshold

typedef std::map<std::string, std::string> Attrs;
typedef std::map<std::string, Attrs> Defaults;

Attrs merged(const Defaults& dflt, const Xml& e) {
    Attrs a;
    auto it = dflt.find(e.tag);
    if (it != dflt.end()) a = it->second;
    for (const auto& kv : e.attrs) a[kv.first] = kv.second;
    return a;
}
const std::string* find(const Attrs& a, const char* k) {
    auto it = a.find(k);
    return it == a.end() ? nullptr : &it->second;
}

struct Geom {
    std::string name;
    int type = RS_GEOM_SPHERE, contype = 1, conaffinity = 1;
    V3 p0, p1, center, half;
    M3 rot;
    double radius = 0, friction = 1, halflen = 0, volume = 0;
    bool fromto = false;
};
struct Joint {
    std::string name;
    int type = RS_JOINT_REVOLUTE;
    V3 axis, pos;
    double lo = 0, hi = -1;
    bool limited = false;
};
struct Body {
    std::string name;
    V3 pos;
    M3 rot;
    std::vector<Joint> joints;
    std::vector<Geom> geoms;
    std::vector<std::unique_ptr<Body>> children;
    bool has_mass = false;
    double mass_override = 0;
};

bool parse_geom(const Xml& e, const Defaults& dflt, bool degrees, Geom& g) {
    const Attrs a = merged(dflt, e);
    if (auto* s = find(a, "name")) g.name = *s;
    const std::string t = find(a, "type") ? *find(a, "type") : "sphere";
    if (t == "plane") return false;
    std::vector<double> size;
    if (auto* s = find(a, "size")) size = numbers(*s);
    const V3 pos = vec3_or(find(a, "pos"), 0, 0, 0);
    if (auto* s = find(a, "quat")) g.rot = quat_wxyz_to_mat(numbers_n(*s, 4));
    else if (auto* s2 = find(a, "axisangle")) g.rot = axisangle_to_mat(numbers_n(*s2, 4), degrees);
    if (auto* s = find(a, "contype")) g.contype = atoi(s->c_str());
    if (auto* s = find(a, "conaffinity")) g.conaffinity = atoi(s->c_str());
    if (auto* s = find(a, "friction")) { auto f = numbers(*s); if (f.empty()) throw std::runtime_error("empty friction"); g.friction = f[0]; }
    auto need = [&](size_t n) { if (size.size() < n) throw std::runtime_error("geom '" + g.name + "': size needs " + std::to_string(n) + " numbers"); };
    if (t == "sphere") {
        need(1);
        g.type = RS_GEOM_SPHERE; g.radius = size[0]; g.p0 = pos; g.p1 = pos;
        g.volume = 4.0 / 3.0 * M_PI * std::pow(g.radius, 3);
    } else if (t == "capsule") {
        need(1);
        g.type = RS_GEOM_CAPSULE; g.radius = size[0];
        double h;
        if (auto* s = find(a, "fromto")) {
            auto ft = numbers_n(*s, 6);
            for (int k = 0; k < 3; k++) { g.p0[k] = ft[k]; g.p1[k] = ft[3 + k]; }
            g.fromto = true;
            h = norm(g.p1 - g.p0);
        } else {
            need(2);
            g.halflen = size[1];
            V3 z; z[2] = -g.halflen;
            g.p0 = pos + mul(g.rot, z);
            z[2] = g.halflen;
            g.p1 = pos + mul(g.rot, z);
            h = 2 * g.halflen;
        }
        g.volume = M_PI * g.radius * g.radius * h + 4.0 / 3.0 * M_PI * std::pow(g.radius, 3);
    } else if (t == "box") {
        if (find(a, "fromto")) throw std::runtime_error("box geom with fromto (geom '" + g.name + "')");
        need(3);
        g.type = RS_GEOM_BOX; g.radius = 0;
        for (int k = 0; k < 3; k++) g.half[k] = size[k];
        g.p0 = pos; g.p1 = pos;
        g.volume = 8.0 * (g.half[0] * g.half[1] * g.half[2]);
    } else {
        throw std::runtime_error("geom type '" + t + "' is not supported (geom '" + g.name + "')");
    }
    g.center = 0.5 * (g.p0 + g.p1);
    return true;
}

Joint parse_joint(const Xml& e, const Defaults& dflt, bool degrees) {
    const Attrs a = merged(dflt, e);
    Joint j;
    if (auto* s = find(a, "name")) j.name = *s;
    const std::string t = find(a, "type") ? *find(a, "type") : "hinge";
    if (t == "hinge") j.type = RS_JOINT_REVOLUTE;
    else if (t == "slide") j.type = RS_JOINT_PRISMATIC;
    else throw std::runtime_error("joint type '" + t + "' is not supported (joint '" + j.name + "')");
    const V3 ax = vec3_or(find(a, "axis"), 0, 0, 1);
    const double n = norm(ax);
    for (int k = 0; k < 3; k++) j.axis[k] = ax[k] / n;
    j.pos = vec3_or(find(a, "pos"), 0, 0, 0);
    const std::string* lim = find(a, "limited");
    const std::string* range = find(a, "range");
    j.limited = lim && *lim == "true" && range;
    j.lo = 0.0; j.hi = -1.0;                      // Bullet's "no limit" convention: lower > upper
    if (j.limited) {
        auto r = numbers_n(*range, 2);
        j.lo = r[0]; j.hi = r[1];
        if (degrees && j.type == RS_JOINT_REVOLUTE) { j.lo = j.lo * (M_PI / 180.0); j.hi = j.hi * (M_PI / 180.0); }
    }
    return j;
}

std::unique_ptr<Body> parse_body(const Xml& e, const Defaults& dflt, bool degrees) {
    std::unique_ptr<Body> b(new Body);
    if (auto* s = e.get("name")) b->name = *s;
    b->pos = vec3_or(e.get("pos"), 0, 0, 0);
    if (auto* s = e.get("quat")) b->rot = quat_wxyz_to_mat(numbers_n(*s, 4));
    for (const Xml& ch : e.kids) {
        if (ch.tag == "joint") b->joints.push_back(parse_joint(ch, dflt, degrees));
        else if (ch.tag == "geom") { Geom g; if (parse_geom(ch, dflt, degrees, g)) b->geoms.push_back(g); }
        else if (ch.tag == "inertial") { if (auto* s = ch.get("mass")) { b->has_mass = true; b->mass_override = strtod(s->c_str(), nullptr); } }
        else if (ch.tag == "body") b->children.push_back(parse_body(ch, dflt, degrees));
    }
    return b;
}

// AABB of one child shape in the link COM frame, as Bullet's getAabb reports it
void geom_aabb(const Geom& g, const V3& com, V3& lo, V3& hi) {
    if (g.type == RS_GEOM_SPHERE) {
        for (int k = 0; k < 3; k++) { const double c = g.p0[k] - com[k]; lo[k] = c - g.radius; hi[k] = c + g.radius; }
    } else if (g.type == RS_GEOM_BOX) {
        const V3 ext = mul_abs(g.rot, g.half);
        for (int k = 0; k < 3; k++) { const double c = g.center[k] - com[k]; lo[k] = c - ext[k]; hi[k] = c + ext[k]; }
    } else if (g.fromto) {
        // btMultiSphereShape: the cached local AABB is exact, getAabb adds one more margin
        for (int k = 0; k < 3; k++) {
            lo[k] = std::fmin(g.p0[k], g.p1[k]) - com[k] - g.radius - MARGIN;
            hi[k] = std::fmax(g.p0[k], g.p1[k]) - com[k] + g.radius + MARGIN;
        }
    } else {
        // btCapsuleShapeZ under a rotated child transform: OBB -> AABB through |R|
        V3 he; he[0] = g.radius; he[1] = g.radius; he[2] = g.radius + g.halflen;
        const V3 ext = mul_abs(g.rot, he);
        for (int k = 0; k < 3; k++) { const double c = g.center[k] - com[k]; lo[k] = c - ext[k]; hi[k] = c + ext[k]; }
    }
}

struct Inertial { double mass = 0; V3 com, inertia; double thresh = BREAKING; };
Inertial body_inertial(const Body& b) {
    Inertial I;
    if (!b.geoms.empty()) {
        const Geom& last = b.geoms.back();
        if (last.type == RS_GEOM_CAPSULE && last.fromto) I.com = last.center;
    }
    if (b.has_mass) I.mass = b.mass_override;
    else { double m = 0; for (const Geom& g : b.geoms) m += 1000.0 * g.volume; I.mass = m; }
    if (b.geoms.empty()) return I;
    V3 lo, hi;
    for (int k = 0; k < 3; k++) { lo[k] = INFINITY; hi[k] = -INFINITY; }
    for (const Geom& g : b.geoms) {
        V3 l, h;
        geom_aabb(g, I.com, l, h);
        for (int k = 0; k < 3; k++) { lo[k] = std::fmin(lo[k], l[k]); hi[k] = std::fmax(hi[k], h[k]); }
    }
    const V3 ext = hi - lo;
    const double lx = ext[0], ly = ext[1], lz = ext[2];
    I.inertia[0] = I.mass / 12.0 * (ly * ly + lz * lz);
    I.inertia[1] = I.mass / 12.0 * (lx * lx + lz * lz);
    I.inertia[2] = I.mass / 12.0 * (lx * lx + ly * ly);
    const double radius = 0.5 * norm(ext);
    const V3 center = 0.5 * (lo + hi);
    I.thresh = BREAKING * (radius + norm(center));
    return I;
}

struct Link {
    int parent = -1, jtype = RS_JOINT_FIXED, limited = 0;
    V3 axis, e, d, inertia;
    double zero_rot[4] = {0, 0, 0, 1};
    double lo = 0, hi = -1, mass = 0, thresh = BREAKING;
    std::string name, jname;
};
struct GeomRef { int link; const Geom* g; V3 com; };

struct Compiler {
    std::vector<Link> links;
    std::vector<GeomRef> geoms;

    // link chain of body b, hanging from real link `parent_link` whose COM sits at parent_com (parent body frame)
    void emit(const Body& b, int parent_link, const V3& parent_com) {
        const Inertial I = body_inertial(b);
        const size_t nj = b.joints.empty() ? 1 : b.joints.size();
        int prev_link = parent_link;
        V3 prev_pivot;
        for (size_t k = 0; k < nj; k++) {
            const bool real = k + 1 == nj;
            const Joint* j = b.joints.empty() ? nullptr : &b.joints[k];
            Link L;
            L.parent = prev_link;
            V3 pj;
            if (!j) { L.jtype = RS_JOINT_FIXED; L.axis[2] = 1.0; }
            else { L.jtype = j->type; L.axis = j->axis; pj = j->pos; L.lo = j->lo; L.hi = j->hi; L.limited = j->limited ? 1 : 0; L.jname = j->name; }
            if (k == 0) {
                L.e = b.pos + mul(b.rot, pj) - parent_com;          // parent frame = parent body frame
                mat_to_quat_xyzw(transpose(b.rot), L.zero_rot);      // parent coords -> this link's coords
            } else {
                L.e = pj - prev_pivot;                               // the previous dummy's COM sits at its pivot
            }
            if (real) { L.d = I.com - pj; L.mass = I.mass; L.inertia = I.inertia; L.thresh = I.thresh; L.name = b.name; }
            else L.name = "dummy_" + L.jname;
            links.push_back(L);
            prev_link = (int)links.size() - 1;
            prev_pivot = pj;
        }
        const int real_idx = (int)links.size() - 1;
        for (const Geom& g : b.geoms) geoms.push_back(GeomRef{real_idx, &g, I.com});
        for (const auto& ch : b.children) emit(*ch, real_idx, I.com);
    }
};

void copy_name(char* dst, const std::string& s) {
    snprintf(dst, RS_NAME_LEN, "%s", s.c_str());
}

void compile(const char* path, double gravity, double timestep, int frame_skip, int has_floor, rs_model_desc* m, rs_model_names* names) {
    FILE* f = fopen(path, "rb");
    if (!f) throw std::runtime_error(std::string("cannot open ") + path);
    std::string src;
    char buf[65536];
    size_t got;
    while ((got = fread(buf, 1, sizeof buf, f)) > 0) src.append(buf, got);
    fclose(f);
    XmlReader rd(src);
    const Xml root = rd.document();

    bool degrees = true;                                   // MuJoCo's default angle unit
    if (const Xml* comp = root.child("compiler")) {
        const std::string* a = comp->get("angle");
        if (a && *a == "radian") degrees = false;
    }
    Defaults dflt;
    if (const Xml* d = root.child("default"))
        for (const Xml& ch : d->kids) { Attrs a; for (const auto& kv : ch.attrs) a[kv.first] = kv.second; dflt[ch.tag] = a; }
    const Xml* wb = root.child("worldbody");
    if (!wb) throw std::runtime_error("no <worldbody>");
    std::vector<std::unique_ptr<Body>> tops;
    for (const Xml& ch : wb->kids) if (ch.tag == "body") tops.push_back(parse_body(ch, dflt, degrees));
    if (tops.empty()) throw std::runtime_error("no top-level body");
    if (tops.size() > 1)
        for (const auto& t : tops)
            if (t->joints.empty()) throw std::runtime_error("several top-level bodies are supported only when each hangs from the world by joints");
    const Body& top = *tops[0];

    memset(m, 0, sizeof *m);
    if (names) memset(names, 0, sizeof *names);
    Compiler C;
    V3 base_pos;
    M3 base_rot;
    std::string base_name;
    if (!top.joints.empty()) {
        // jointed root(s): a massless fixed anchor at the world origin; every top-level body hangs from it by its own joints
        m->fixed_base = 1;
        m->base_mass = 0.0;
        base_name = "world";
        m->break_thresh[0] = BREAKING;
        for (const auto& t : tops) C.emit(*t, -1, V3());
    } else {
        m->fixed_base = 0;
        const Inertial I = body_inertial(top);
        m->base_mass = I.mass;
        for (int k = 0; k < 3; k++) m->base_inertia[k] = I.inertia[k];
        base_name = top.name;
        m->break_thresh[0] = I.thresh;
        for (const Geom& g : top.geoms) C.geoms.push_back(GeomRef{-1, &g, I.com});
        for (const auto& ch : top.children) C.emit(*ch, -1, I.com);
        base_pos = top.pos + mul(top.rot, I.com);
        base_rot = top.rot;
    }
    if (C.links.size() > RS_MAX_LINKS) throw std::runtime_error("too many links (" + std::to_string(C.links.size()) + " > RS_MAX_LINKS)");
    if (C.geoms.size() > RS_MAX_GEOMS) throw std::runtime_error("too many geoms (" + std::to_string(C.geoms.size()) + " > RS_MAX_GEOMS)");
    m->n_links = (int)C.links.size();
    int ndof = 0;
    for (int i = 0; i < m->n_links; i++) {
        const Link& L = C.links[i];
        m->parent[i] = L.parent;
        m->jtype[i] = L.jtype;
        m->dof_idx[i] = L.jtype == RS_JOINT_FIXED ? -1 : ndof++;
        m->has_limit[i] = L.limited;
        for (int k = 0; k < 3; k++) { m->axis[i][k] = L.axis[k]; m->e_vec[i][k] = L.e[k]; m->d_vec[i][k] = L.d[k]; m->inertia[i][k] = L.inertia[k]; }
        for (int k = 0; k < 4; k++) m->zero_rot[i][k] = L.zero_rot[k];
        m->mass[i] = L.mass;
        m->lim_lo[i] = L.lo; m->lim_hi[i] = L.hi;
        m->break_thresh[i + 1] = L.thresh;
        if (names) { copy_name(names->link_names[i], L.name); copy_name(names->joint_names[i], L.jname); }
    }
    m->n_dof = ndof;
    if (names) copy_name(names->base_name, base_name);

    m->n_geoms = (int)C.geoms.size();
    const int floor_contype = 1, floor_conaff = 1;          // ground_plane.xml:3
    for (int gi = 0; gi < m->n_geoms; gi++) {
        const GeomRef& r = C.geoms[gi];
        const Geom& g = *r.g;
        m->g_link[gi] = r.link;
        m->g_type[gi] = g.type;
        for (int k = 0; k < 3; k++) {
            m->g_p0[gi][k] = g.p0[k] - r.com[k];
            m->g_p1[gi][k] = g.type == RS_GEOM_BOX ? g.half[k] : g.p1[k] - r.com[k];
        }
        if (g.type == RS_GEOM_BOX) mat_to_quat_xyzw(g.rot, m->g_quat[gi]);
        else { m->g_quat[gi][0] = m->g_quat[gi][1] = m->g_quat[gi][2] = 0.0; m->g_quat[gi][3] = 1.0; }
        m->g_radius[gi] = g.radius;
        m->g_friction[gi] = g.friction;
        m->g_floor[gi] = (has_floor && ((g.contype & floor_conaff) || (floor_contype & g.conaffinity))) ? 1 : 0;
        if (names) copy_name(names->geom_names[gi], g.name);
    }
    m->has_floor = has_floor ? 1 : 0;
    m->floor_friction = 0.8;

    // self-collision pairs: URDF_USE_SELF_COLLISION | URDF_USE_SELF_COLLISION_EXCLUDE_ALL_PARENTS, MuJoCo OR filter
    std::vector<std::set<int>> anc(C.links.size());
    for (size_t li = 0; li < C.links.size(); li++) {
        int p = (int)li;
        while (p != -1) { p = C.links[p].parent; anc[li].insert(p); }
    }
    auto is_anc = [&](int a, int of) { return of >= 0 && anc[of].count(a) != 0; };   // a is an ancestor of `of`
    int np = 0;
    for (int a = 0; a < m->n_geoms; a++)
        for (int b = a + 1; b < m->n_geoms; b++) {
            const int la = C.geoms[a].link, lb = C.geoms[b].link;
            const Geom& ga = *C.geoms[a].g; const Geom& gb = *C.geoms[b].g;
            if (la == lb || is_anc(la, lb) || is_anc(lb, la)) continue;
            if (!((ga.contype & gb.conaffinity) || (gb.contype & ga.conaffinity))) continue;
            if (ga.type == RS_GEOM_BOX || gb.type == RS_GEOM_BOX) continue;     // box pairs need GJK/EPA: only box-vs-plane is built
            if (np >= RS_MAX_PAIRS) throw std::runtime_error("too many self-collision pairs (> RS_MAX_PAIRS)");
            m->pair_a[np] = a; m->pair_b[np] = b; np++;
        }
    m->n_pairs = np;

    // world / solver constants; gravity and timestep squeezed through float32 like World(float, float) (python-binding.cpp:297)
    m->gravity = (double)(float)gravity;
    m->timestep = (double)(float)timestep;
    m->frame_skip = frame_skip;
    m->n_iter = 5;                  // physics-bullet.cpp:370
    m->erp2 = 0.9;                  // physics-bullet.cpp:371 (contactERP)
    m->erp = 0.2;                   // btContactSolverInfo m_erp
    m->split_thresh = -0.04;        // m_splitImpulsePenetrationThreshold
    m->c_erp = m->erp; m->c_erp2 = m->erp2; m->c_split_thresh = m->split_thresh;
    m->lin_damping = 0.04; m->ang_damping = 0.04;      // btMultiBody constructor
    m->max_coord_vel = 100.0;
    m->max_limit_impulse = 100.0;   // btMultiBodyConstraint m_maxAppliedImpulse
    m->use_gyro = 1.0;
    for (int k = 0; k < 3; k++) m->reset_base_pos[k] = base_pos[k];
    mat_to_quat_xyzw(base_rot, m->reset_base_quat);
    m->alive_rule = RS_ALIVE_ALWAYS;
    m->body_link = -1;
    m->task_link = m->task_link2 = -1;       // no task link until a task spec names one (a link-less body has no link 0)
    m->initial_z = -1.0;
    m->max_episode_steps = 1000;
    m->max_contacts = 16;                               // live contact points per env (envspec.py's default; Hopper-size models may use 8)
    m->dt_env = timestep * frame_skip;                  // scene.dt, a Python double (scene_abstract.py:22)
}

}  // namespace

extern "C" int rs_model_compile(const char* mjcf_path, double gravity, double timestep, int frame_skip, int has_floor,
                                rs_model_desc* out, rs_model_names* names) {
    if (!mjcf_path || !out) { return rs_set_error("rs_model_compile: null argument"); }
    try {
        compile(mjcf_path, gravity, timestep, frame_skip, has_floor, out, names);
    } catch (const std::exception& e) {
        return rs_set_error((std::string("rs_model_compile: ") + e.what()).c_str());
    }
    return 0;
}
