// scene_file.cpp -- the reference's scene text grammar + Wavefront OBJ -> flattened world-space arrays.
//
// Replaces Scene::Scene(filename) (scene.cpp:96-131), loadMaterial (:376-433), loadModel (:222-286),
// loadCamera (:288-355), Resource::loadOBJMesh (:23-55) and the flattening loop of buildDevData (:159-190).
// The arithmetic of the instance transform follows glm 0.9.6.3 (vendored by the reference) operation by
// operation so that the flattened arrays are bit-identical: gtc/matrix_transform.inl:40-134 (translate /
// rotate / scale), detail/type_mat4x4.inl:37-92 (inverse), :592-628 (mat4*vec4), :686-704 (mat4*mat4).
// OBJ numbers are parsed with tinyobjloader v2.0.0's own decimal parser (tiny_obj_loader.h:866-1000,
// restated below), not strtod, because the two differ in the last bit for long mantissas.
#include <ctype.h>
#include <string.h>

#include <fstream>
#include <limits>
#include <map>
#include <sstream>

#include "scene_host.h"

namespace rs {

namespace {

// ---------------------------------------------------------------------------- text helpers
// utilities.cpp:65-95 safeGetline: accepts \n, \r\n, \r and a last line without terminator
bool safeGetline(std::istream& is, std::string& t) {
    t.clear();
    std::streambuf* sb = is.rdbuf();
    if (!is.good()) return false;
    for (;;) {
        int c = sb->sbumpc();
        switch (c) {
        case '\n': return true;
        case '\r':
            if (sb->sgetc() == '\n') sb->sbumpc();
            return true;
        case EOF:
            is.setstate(t.empty() ? (std::ios::eofbit | std::ios::failbit) : std::ios::eofbit);
            return !t.empty();
        default: t += (char)c;
        }
    }
}
std::vector<std::string> tokenize(const std::string& s) {      // utilities.cpp:57-63
    std::istringstream ss(s);
    std::vector<std::string> out;
    std::string tok;
    while (ss >> tok) out.push_back(tok);
    return out;
}

// ---------------------------------------------------------------------------- glm-exact 4x4 helpers
struct V4 { float x, y, z, w; };
inline V4 operator*(V4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline V4 operator*(V4 a, V4 b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
inline V4 operator+(V4 a, V4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline V4 operator-(V4 a, V4 b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
struct M4 { V4 c[4]; };   // column-major, c[col]
inline float& at(M4& m, int col, int row) { return (&m.c[col].x)[row]; }
inline float at(const M4& m, int col, int row) { return (&m.c[col].x)[row]; }
M4 identity() { M4 m; m.c[0] = {1, 0, 0, 0}; m.c[1] = {0, 1, 0, 0}; m.c[2] = {0, 0, 1, 0}; m.c[3] = {0, 0, 0, 1}; return m; }
M4 mul(const M4& a, const M4& b) {                              // type_mat4x4.inl:686-704
    M4 r;
    for (int j = 0; j < 4; j++) {
        const V4 bj = b.c[j];
        r.c[j] = a.c[0] * bj.x + a.c[1] * bj.y + a.c[2] * bj.z + a.c[3] * bj.w;
    }
    return r;
}
M4 translate(const M4& m, f3 v) {                               // matrix_transform.inl:40-49
    M4 r = m;
    r.c[3] = m.c[0] * v.x + m.c[1] * v.y + m.c[2] * v.z + m.c[3];
    return r;
}
M4 rotate(const M4& m, float angle, f3 v) {                     // matrix_transform.inl:52-85
    const float c = cosf(angle), s = sinf(angle);
    f3 axis = normalize(v);
    f3 temp = axis * (1.f - c);
    float R[3][3];
    R[0][0] = c + temp.x * axis.x;
    R[0][1] = 0 + temp.x * axis.y + s * axis.z;
    R[0][2] = 0 + temp.x * axis.z - s * axis.y;
    R[1][0] = 0 + temp.y * axis.x - s * axis.z;
    R[1][1] = c + temp.y * axis.y;
    R[1][2] = 0 + temp.y * axis.z + s * axis.x;
    R[2][0] = 0 + temp.z * axis.x + s * axis.y;
    R[2][1] = 0 + temp.z * axis.y - s * axis.x;
    R[2][2] = c + temp.z * axis.z;
    M4 r;
    r.c[0] = m.c[0] * R[0][0] + m.c[1] * R[0][1] + m.c[2] * R[0][2];
    r.c[1] = m.c[0] * R[1][0] + m.c[1] * R[1][1] + m.c[2] * R[1][2];
    r.c[2] = m.c[0] * R[2][0] + m.c[1] * R[2][1] + m.c[2] * R[2][2];
    r.c[3] = m.c[3];
    return r;
}
M4 scale(const M4& m, f3 v) {                                   // matrix_transform.inl:122-134
    M4 r;
    r.c[0] = m.c[0] * v.x; r.c[1] = m.c[1] * v.y; r.c[2] = m.c[2] * v.z; r.c[3] = m.c[3];
    return r;
}
M4 inverse(const M4& m) {                                       // type_mat4x4.inl:37-92; m[col][row]
#define M(cc, rr) at(m, cc, rr)
    float Coef00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3);
    float Coef02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3);
    float Coef03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3);
    float Coef04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
    float Coef06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3);
    float Coef07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
    float Coef08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2);
    float Coef10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2);
    float Coef11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2);
    float Coef12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3);
    float Coef14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3);
    float Coef15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
    float Coef16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2);
    float Coef18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2);
    float Coef19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2);
    float Coef20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1);
    float Coef22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1);
    float Coef23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);
    V4 Fac0 = {Coef00, Coef00, Coef02, Coef03}, Fac1 = {Coef04, Coef04, Coef06, Coef07}, Fac2 = {Coef08, Coef08, Coef10, Coef11};
    V4 Fac3 = {Coef12, Coef12, Coef14, Coef15}, Fac4 = {Coef16, Coef16, Coef18, Coef19}, Fac5 = {Coef20, Coef20, Coef22, Coef23};
    V4 Vec0 = {M(1, 0), M(0, 0), M(0, 0), M(0, 0)}, Vec1 = {M(1, 1), M(0, 1), M(0, 1), M(0, 1)};
    V4 Vec2 = {M(1, 2), M(0, 2), M(0, 2), M(0, 2)}, Vec3 = {M(1, 3), M(0, 3), M(0, 3), M(0, 3)};
    V4 Inv0 = Vec1 * Fac0 - Vec2 * Fac1 + Vec3 * Fac2;
    V4 Inv1 = Vec0 * Fac0 - Vec2 * Fac3 + Vec3 * Fac4;
    V4 Inv2 = Vec0 * Fac1 - Vec1 * Fac3 + Vec3 * Fac5;
    V4 Inv3 = Vec0 * Fac2 - Vec1 * Fac4 + Vec2 * Fac5;
    V4 SignA = {+1, -1, +1, -1}, SignB = {-1, +1, -1, +1};
    M4 Inv;
    Inv.c[0] = Inv0 * SignA; Inv.c[1] = Inv1 * SignB; Inv.c[2] = Inv2 * SignA; Inv.c[3] = Inv3 * SignB;
    V4 Row0 = {Inv.c[0].x, Inv.c[1].x, Inv.c[2].x, Inv.c[3].x};
    V4 Dot0 = m.c[0] * Row0;
    float Dot1 = (Dot0.x + Dot0.y) + (Dot0.z + Dot0.w);
    float ood = 1.f / Dot1;
    for (int i = 0; i < 4; i++) Inv.c[i] = Inv.c[i] * ood;
#undef M
    return Inv;
}
V4 mulVec(const M4& m, V4 v) {                                  // type_mat4x4.inl:617-628
    V4 Mul0 = m.c[0] * v.x, Mul1 = m.c[1] * v.y;
    V4 Add0 = Mul0 + Mul1;
    V4 Mul2 = m.c[2] * v.z, Mul3 = m.c[3] * v.w;
    V4 Add1 = Mul2 + Mul3;
    return Add0 + Add1;
}

// ---------------------------------------------------------------------------- OBJ
// tinyobjloader v2.0.0 tryParseDouble (tiny_obj_loader.h:866-1000), restated
bool tinyParseDouble(const char* s, const char* s_end, double* result) {
    if (s >= s_end) return false;
    double mantissa = 0.0;
    int exponent = 0;
    char sign = '+', exp_sign = '+';
    const char* curr = s;
    int read = 0;
    bool end_not_reached = false, leading_dot = false;
    if (*curr == '+' || *curr == '-') {
        sign = *curr; curr++;
        if (curr != s_end && *curr == '.') leading_dot = true;
    } else if (isdigit((unsigned char)*curr)) {
    } else if (*curr == '.') {
        leading_dot = true;
    } else return false;
    end_not_reached = (curr != s_end);
    if (!leading_dot) {
        while (end_not_reached && isdigit((unsigned char)*curr)) {
            mantissa *= 10;
            mantissa += (int)(*curr - 0x30);
            curr++; read++;
            end_not_reached = (curr != s_end);
        }
        if (read == 0) return false;
    }
    bool assemble = !end_not_reached;
    if (!assemble) {
        if (*curr == '.') {
            curr++; read = 1;
            end_not_reached = (curr != s_end);
            static const double pow_lut[] = {1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001};
            while (end_not_reached && isdigit((unsigned char)*curr)) {
                mantissa += (int)(*curr - 0x30) * (read < 8 ? pow_lut[read] : pow(10.0, -read));
                read++; curr++;
                end_not_reached = (curr != s_end);
            }
        } else if (*curr == 'e' || *curr == 'E') {
        } else assemble = true;
    }
    if (!assemble && end_not_reached && (*curr == 'e' || *curr == 'E')) {
        curr++;
        end_not_reached = (curr != s_end);
        if (end_not_reached && (*curr == '+' || *curr == '-')) { exp_sign = *curr; curr++; }
        else if (end_not_reached && isdigit((unsigned char)*curr)) {
        } else return false;
        read = 0;
        end_not_reached = (curr != s_end);
        while (end_not_reached && isdigit((unsigned char)*curr)) {
            if (exponent > 2147483647 / 10) return false;
            exponent *= 10;
            exponent += (int)(*curr - 0x30);
            curr++; read++;
            end_not_reached = (curr != s_end);
        }
        exponent *= (exp_sign == '+' ? 1 : -1);
        if (read == 0) return false;
    }
    *result = (sign == '+' ? 1 : -1) * (exponent ? ldexp(mantissa * pow(5.0, exponent), exponent) : mantissa);
    return true;
}
float tinyParseReal(const char*& tok) {                         // parseReal, tiny_obj_loader.h:1008-1017
    tok += strspn(tok, " \t");
    const char* end = tok + strcspn(tok, " \t\r");
    double val = 0.0;
    tinyParseDouble(tok, end, &val);
    tok = end;
    return (float)val;
}

struct Mesh { std::vector<f3> v, n; std::vector<float> t; };    // de-indexed, 3 entries per triangle (scene.cpp:41-50)

bool fixIndex(int idx, int n, int* out) {                       // tiny_obj_loader.h fixIndex
    if (idx > 0) { *out = idx - 1; return true; }
    if (idx == 0) return false;
    *out = n + idx;
    return true;
}


struct Idx { int v, vt, vn; };     // one corner of an OBJ face: position / texcoord / normal index (0-based, -1 = absent)

// Polygons with five or more corners, as tinyobj's built-in ear clipping splits them (tiny_obj_loader.h:1536-1821,
// v2.0 without TINYOBJLOADER_USE_MAPBOX_EARCUT): project on the two axes orthogonal-ish to the first non-degenerate
// corner's normal, then repeatedly cut the corner at `guess` when it is convex w.r.t. the (origin-based) area sign and
// no other corner lies inside it (pnpoly, :1388-1398); give up after a full fruitless round.  The emitted corner order
// fixes the flattened primitive order, hence everything downstream (BVH, light list): it has to be reproduced exactly.
static bool pointInTri(const float* vx, const float* vy, float tx, float ty) {
    bool c = false;
    for (int i = 0, j = 2; i < 3; j = i++)
        if (((vy[i] > ty) != (vy[j] > ty)) && (tx < (vx[j] - vx[i]) * (ty - vy[i]) / (vy[j] - vy[i]) + vx[i])) c = !c;
    return c;
}
static bool earClip(const std::vector<Idx>& face, const std::vector<float>& V, std::vector<Idx>& tri) {
    const size_t n0 = face.size();
    for (const Idx& ix : face)
        if (ix.v < 0 || 3 * (size_t)ix.v + 2 >= V.size()) return false;
    size_t axes[2] = {1, 2};
    for (size_t k = 0; k < n0; ++k) {
        const float* p0 = &V[3 * face[k % n0].v]; const float* p1 = &V[3 * face[(k + 1) % n0].v]; const float* p2 = &V[3 * face[(k + 2) % n0].v];
        float e0x = p1[0] - p0[0], e0y = p1[1] - p0[1], e0z = p1[2] - p0[2];
        float e1x = p2[0] - p1[0], e1y = p2[1] - p1[1], e1z = p2[2] - p1[2];
        float cx = fabsf(e0y * e1z - e0z * e1y), cy = fabsf(e0z * e1x - e0x * e1z), cz = fabsf(e0x * e1y - e0y * e1x);
        const float eps = std::numeric_limits<float>::epsilon();
        if (cx > eps || cy > eps || cz > eps) {
            if (!(cx > cy && cx > cz)) {
                axes[0] = 0;
                if (cz > cx && cz > cy) axes[1] = 1;
            }
            break;
        }
    }
    std::vector<Idx> rest = face;
    size_t guess = 0, budget = n0, prevCount = rest.size();
    while (rest.size() > 3 && budget > 0) {
        const size_t n = rest.size();
        if (guess >= n) guess -= n;
        if (prevCount != n) { prevCount = n; budget = n; }
        else budget--;
        Idx ind[3];
        float vx[3], vy[3];
        for (int k = 0; k < 3; k++) {
            ind[k] = rest[(guess + k) % n];
            vx[k] = V[3 * ind[k].v + axes[0]];
            vy[k] = V[3 * ind[k].v + axes[1]];
        }
        float e0x = vx[1] - vx[0], e0y = vy[1] - vy[0], e1x = vx[2] - vx[1], e1y = vy[2] - vy[1];
        float cross = e0x * e1y - e0y * e1x;
        float area = (vx[0] * vy[1] - vy[0] * vx[1]) * 0.5f;
        if (cross * area < 0.0f) { guess += 1; continue; }
        bool overlap = false;
        for (size_t other = 3; other < n; ++other) {
            const Idx& o = rest[(guess + other) % n];
            if (pointInTri(vx, vy, V[3 * o.v + axes[0]], V[3 * o.v + axes[1]])) { overlap = true; break; }
        }
        if (overlap) { guess += 1; continue; }
        tri.push_back(ind[0]); tri.push_back(ind[1]); tri.push_back(ind[2]);
        rest.erase(rest.begin() + (guess + 1) % n);
    }
    if (rest.size() == 3) { tri.push_back(rest[0]); tri.push_back(rest[1]); tri.push_back(rest[2]); }
    return true;
}

bool loadOBJ(const std::string& path, Mesh& mesh, std::string& err) {
    std::ifstream in(path.c_str());
    if (!in.is_open()) { err = "cannot open OBJ file " + path; return false; }
    std::vector<float> V, N, TC;
    std::string line;
    while (safeGetline(in, line)) {
        const char* tok = line.c_str();
        tok += strspn(tok, " \t");
        if (tok[0] == '\0' || tok[0] == '#') continue;
        if (tok[0] == 'v' && (tok[1] == ' ' || tok[1] == '\t')) {
            tok += 2;
            float x = tinyParseReal(tok), y = tinyParseReal(tok), z = tinyParseReal(tok);
            V.push_back(x); V.push_back(y); V.push_back(z);
        } else if (tok[0] == 'v' && tok[1] == 'n' && (tok[2] == ' ' || tok[2] == '\t')) {
            tok += 3;
            float x = tinyParseReal(tok), y = tinyParseReal(tok), z = tinyParseReal(tok);
            N.push_back(x); N.push_back(y); N.push_back(z);
        } else if (tok[0] == 'v' && tok[1] == 't' && (tok[2] == ' ' || tok[2] == '\t')) {
            tok += 3;
            float x = tinyParseReal(tok), y = tinyParseReal(tok);
            TC.push_back(x); TC.push_back(y);
        } else if (tok[0] == 'f' && (tok[1] == ' ' || tok[1] == '\t')) {
            tok += 2;
            std::vector<Idx> face;
            for (;;) {
                tok += strspn(tok, " \t");
                if (*tok == '\0' || *tok == '\r' || *tok == '\n') break;
                Idx ix = {-1, -1, -1};
                int raw = atoi(tok);
                if (!fixIndex(raw, (int)V.size() / 3, &ix.v)) { err = "OBJ face index 0 in " + path; return false; }
                tok += strcspn(tok, "/ \t\r");
                if (*tok == '/') {
                    tok++;
                    if (*tok == '/') {
                        tok++;
                        if (!fixIndex(atoi(tok), (int)N.size() / 3, &ix.vn)) { err = "OBJ normal index 0 in " + path; return false; }
                        tok += strcspn(tok, "/ \t\r");
                    } else {
                        if (!fixIndex(atoi(tok), (int)TC.size() / 2, &ix.vt)) { err = "OBJ texcoord index 0 in " + path; return false; }
                        tok += strcspn(tok, "/ \t\r");
                        if (*tok == '/') {
                            tok++;
                            if (!fixIndex(atoi(tok), (int)N.size() / 3, &ix.vn)) { err = "OBJ normal index 0 in " + path; return false; }
                            tok += strcspn(tok, "/ \t\r");
                        }
                    }
                }
                face.push_back(ix);
            }
            std::vector<Idx> tri;
            if (face.size() == 3) tri = face;
            else if (face.size() == 4) {                         // tiny_obj_loader.h:1429-1535: split along the shorter diagonal
                const float* a = &V[3 * face[0].v]; const float* b = &V[3 * face[1].v];
                const float* c = &V[3 * face[2].v]; const float* d = &V[3 * face[3].v];
                float e02x = c[0] - a[0], e02y = c[1] - a[1], e02z = c[2] - a[2];
                float e13x = d[0] - b[0], e13y = d[1] - b[1], e13z = d[2] - b[2];
                float sqr02 = e02x * e02x + e02y * e02y + e02z * e02z;
                float sqr13 = e13x * e13x + e13y * e13y + e13z * e13z;
                if (sqr02 < sqr13) tri = {face[0], face[1], face[2], face[0], face[2], face[3]};
                else tri = {face[0], face[1], face[3], face[1], face[2], face[3]};
            } else if (face.size() < 3) continue;
            else if (!earClip(face, V, tri)) { err = "OBJ polygon references a vertex that does not exist: " + path; return false; }
            for (const Idx& ix : tri) {
                if (ix.v < 0 || 3 * ix.v + 2 >= (int)V.size()) { err = "OBJ vertex index out of range in " + path; return false; }
                if (ix.vn < 0 || 3 * ix.vn + 2 >= (int)N.size()) { err = "OBJ faces need normals (scene.cpp:44): " + path; return false; }
                mesh.v.push_back(mk3(V[3 * ix.v], V[3 * ix.v + 1], V[3 * ix.v + 2]));
                mesh.n.push_back(mk3(N[3 * ix.vn], N[3 * ix.vn + 1], N[3 * ix.vn + 2]));
                if (!TC.empty() && ix.vt >= 0 && 2 * ix.vt + 1 < (int)TC.size()) { mesh.t.push_back(TC[2 * ix.vt]); mesh.t.push_back(TC[2 * ix.vt + 1]); }
                else { mesh.t.push_back(0.f); mesh.t.push_back(0.f); }
            }
        }
    }
    return true;
}

struct Instance { f3 translation, rotation, scale; int materialId; const Mesh* mesh; };

}  // namespace

bool loadSceneFile(const std::string& path, HostScene& hs, RstrCamera& cam, std::string& err) {
    std::ifstream in(path.c_str());
    if (!in.is_open()) { err = "cannot open scene file " + path; return false; }
    std::map<std::string, int> materialMap;
    std::map<std::string, Mesh> meshPool;
    std::vector<Instance> instances;
    const std::map<std::string, int> typeMap = {{"Lambertian", 0}, {"MetallicWorkflow", 1}, {"Dielectric", 2}, {"Light", 4}};
    auto stof3 = [](const std::vector<std::string>& t) { return mk3(std::stof(t.at(1)), std::stof(t.at(2)), std::stof(t.at(3))); };
    // Scene::addTexture / Resource::loadTexture (scene.cpp:362-375): one Image per file NAME (the first load wins, so
    // its vertical-flip setting too), texture ids in order of first use
    std::map<std::string, int> textureMap;
    hs.textures.clear(); hs.envMapTexId = -1;
    bool flipOnLoad = true;                                                           // scene.cpp:98
    std::string texErr;
    auto addTexture = [&](const std::string& filename) -> int {
        auto it = textureMap.find(filename);
        if (it != textureMap.end()) return it->second;
        std::string resolved = filename;
        if (!std::ifstream(resolved.c_str()).good()) {                                // also try relative to the scene file
            size_t slash = path.find_last_of('/');
            if (slash != std::string::npos) resolved = path.substr(0, slash + 1) + filename;
        }
        HostTexture tex;
        if (!loadImageRGB(resolved, flipOnLoad, tex, texErr)) return -3;
        int id = (int)hs.textures.size();
        hs.textures.push_back(std::move(tex));
        textureMap[filename] = id;
        return id;
    };
    std::string line;
    try {
        while (in.good()) {
            safeGetline(in, line);
            if (line.empty()) continue;
            std::vector<std::string> tokens = tokenize(line);
            if (tokens.empty()) continue;
            if (tokens[0] == "Material") {                                            // scene.cpp:376-433
                RstrMaterial m;
                m.type = 0; m.baseColor[0] = m.baseColor[1] = m.baseColor[2] = .9f; m.metallic = 0.f; m.roughness = 1.f; m.ior = 1.5f;
                m.baseColorMapId = m.metallicMapId = m.roughnessMapId = m.normalMapId = -1;
                for (int i = 0; i < 6; i++) {
                    safeGetline(in, line);
                    auto t = tokenize(line);
                    if (t.empty()) continue;
                    if (t[0] == "Type") {
                        auto it = typeMap.find(t.at(1));
                        m.type = it == typeMap.end() ? 0 : it->second;                // std::map::operator[] default-inserts 0
                    } else if (t[0] == "BaseColor") {
                        if (t.size() > 2) { f3 c = stof3(t); m.baseColor[0] = c.x; m.baseColor[1] = c.y; m.baseColor[2] = c.z; }
                        else if (t.at(1) == "Procedural") m.baseColorMapId = -2;         // ProceduralTexId, material.h:13
                        else if ((m.baseColorMapId = addTexture(t[1])) == -3) { err = texErr; return false; }
                    } else if (t[0] == "Metallic") {
                        if (isdigit((unsigned char)t.at(1).back())) m.metallic = std::stof(t[1]);
                        else if ((m.metallicMapId = addTexture(t[1])) == -3) { err = texErr; return false; }
                    } else if (t[0] == "Roughness") {
                        if (isdigit((unsigned char)t.at(1).back())) m.roughness = std::stof(t[1]);
                        else if ((m.roughnessMapId = addTexture(t[1])) == -3) { err = texErr; return false; }
                    } else if (t[0] == "Ior") {
                        m.ior = std::stof(t.at(1));
                    } else if (t[0] == "NormalMap") {
                        if (t.at(1) != "Null" && (m.normalMapId = addTexture(t[1])) == -3) { err = texErr; return false; }
                    }
                }
                materialMap[tokens.at(1)] = (int)hs.materials.size();
                hs.materials.push_back(m);
            } else if (tokens[0] == "Object") {                                       // scene.cpp:222-286
                Instance inst;
                inst.translation = inst.rotation = inst.scale = mk3(0.f);             // glm 0.9.6 zero-initialises vec3
                inst.materialId = 0;
                safeGetline(in, line);
                std::string filename = line;
                if (filename.find(".obj") == std::string::npos) { err = "only .obj meshes are supported (loadGLTFMesh is a stub, scene.cpp:57-63): " + filename; return false; }
                auto it = meshPool.find(filename);
                if (it == meshPool.end()) {
                    Mesh mesh;
                    std::string resolved = filename;
                    if (!std::ifstream(resolved.c_str()).good()) {                    // also try relative to the scene file
                        size_t slash = path.find_last_of('/');
                        if (slash != std::string::npos) resolved = path.substr(0, slash + 1) + filename;
                    }
                    if (!loadOBJ(resolved, mesh, err)) return false;
                    it = meshPool.emplace(filename, std::move(mesh)).first;
                }
                inst.mesh = &it->second;
                safeGetline(in, line);
                if (!line.empty() && in.good()) {
                    auto t = tokenize(line);
                    if (t.at(1) == "Null") {
                        RstrMaterial m;
                        m.type = 0; m.baseColor[0] = m.baseColor[1] = m.baseColor[2] = .9f; m.metallic = 0.f; m.roughness = 1.f; m.ior = 1.5f;
                        m.baseColorMapId = m.metallicMapId = m.roughnessMapId = m.normalMapId = -1;
                        inst.materialId = (int)hs.materials.size();
                        hs.materials.push_back(m);
                    } else {
                        auto mi = materialMap.find(t[1]);
                        if (mi == materialMap.end()) { err = "Material " + t[1] + " doesn't exist (scene.cpp:251-254)"; return false; }
                        inst.materialId = mi->second;
                    }
                }
                safeGetline(in, line);
                while (!line.empty() && in.good()) {
                    auto t = tokenize(line);
                    if (!t.empty()) {
                        if (t[0] == "Translate") inst.translation = stof3(t);
                        else if (t[0] == "Rotate") inst.rotation = stof3(t);
                        else if (t[0] == "Scale") inst.scale = stof3(t);
                    }
                    safeGetline(in, line);
                }
                instances.push_back(inst);
            } else if (tokens[0] == "Camera") {                                       // scene.cpp:288-355
                float fovy = 0.f;
                for (int i = 0; i < 8; i++) {
                    safeGetline(in, line);
                    auto t = tokenize(line);
                    if (t.empty()) continue;
                    if (t[0] == "Resolution") { cam.resolution[0] = std::stoi(t.at(1)); cam.resolution[1] = std::stoi(t.at(2)); }
                    else if (t[0] == "FovY") fovy = std::stof(t.at(1));
                    else if (t[0] == "LensRadius") cam.lensRadius = std::stof(t.at(1));
                    else if (t[0] == "FocalDist") cam.focalDist = std::stof(t.at(1));
                }
                safeGetline(in, line);
                while (!line.empty() && in.good()) {
                    auto t = tokenize(line);
                    if (!t.empty()) {
                        f3 v = mk3(0.f);
                        if (t[0] == "Eye" || t[0] == "Rotation" || t[0] == "Up") v = stof3(t);
                        if (t[0] == "Eye") memcpy(cam.position, &v, 12);
                        else if (t[0] == "Rotation") memcpy(cam.rotation, &v, 12);
                        else if (t[0] == "Up") memcpy(cam.up, &v, 12);
                    }
                    safeGetline(in, line);
                }
                float yscaled = tanf(fovy * (RS_PI / 180));                           // scene.cpp:344-348
                float xscaled = (yscaled * cam.resolution[0]) / cam.resolution[1];
                float fovx = (atanf(xscaled) * 180) / RS_PI;
                cam.fov[0] = fovx; cam.fov[1] = fovy;
                cam.tanFovY = tanf(radians(fovy * 0.5f));
                cameraUpdate(cam);
            } else if (tokens[0] == "EnvMap") {
                if (tokens.at(1) != "Null") {                                         // scene.cpp:122-128: loaded unflipped
                    flipOnLoad = false;
                    hs.envMapTexId = addTexture(tokens[1]);
                    flipOnLoad = true;
                    if (hs.envMapTexId == -3) { err = texErr; return false; }
                }
            }
        }
    } catch (const std::exception& e) {
        err = std::string("scene file parse error near '") + line + "': " + e.what();
        return false;
    }
    // ---- flatten (scene.cpp:159-190) ----
    hs.vertices.clear(); hs.normals.clear(); hs.texcoords.clear(); hs.materialIds.clear(); hs.lightPowerFromFile.clear();
    for (const Instance& inst : instances) {
        const RstrMaterial& material = hs.materials[inst.materialId];
        const float powerUnitArea = luminance(mk3(material.baseColor[0], material.baseColor[1], material.baseColor[2])) * 2.f * RS_GLM_PI;
        // Math::buildTransformationMatrix (mathUtil.cpp:13-19)
        M4 translationMat = translate(identity(), inst.translation);
        M4 rotationMat = rotate(identity(), inst.rotation.x * RS_PI / 180.f, mk3(1.f, 0.f, 0.f));
        rotationMat = mul(rotationMat, rotate(identity(), inst.rotation.y * RS_PI / 180.f, mk3(0.f, 1.f, 0.f)));
        rotationMat = mul(rotationMat, rotate(identity(), inst.rotation.z * RS_PI / 180.f, mk3(0.f, 0.f, 1.f)));
        M4 scaleMat = scale(identity(), inst.scale);
        M4 transform = mul(mul(translationMat, rotationMat), scaleMat);
        M4 inv = inverse(transform);                                                  // scene.cpp:282-283
        // normalMat = transpose(mat3(inv)): normalMat[col][row] = inv[row][col]
        const Mesh& mesh = *inst.mesh;
        for (size_t i = 0; i < mesh.v.size(); i++) {
            V4 p = mulVec(transform, V4{mesh.v[i].x, mesh.v[i].y, mesh.v[i].z, 1.f});
            hs.vertices.push_back(mk3(p.x, p.y, p.z));
            f3 n = mesh.n[i];
            // mat3 * vec3 (type_mat3x3.inl:487): result[r] = nm[0][r]*n.x + nm[1][r]*n.y + nm[2][r]*n.z, nm[c][r] = inv[r][c]
            f3 q = mk3(at(inv, 0, 0) * n.x + at(inv, 0, 1) * n.y + at(inv, 0, 2) * n.z,
                       at(inv, 1, 0) * n.x + at(inv, 1, 1) * n.y + at(inv, 1, 2) * n.z,
                       at(inv, 2, 0) * n.x + at(inv, 2, 1) * n.y + at(inv, 2, 2) * n.z);
            hs.normals.push_back(normalize(q));
            hs.texcoords.push_back(mesh.t[2 * i]); hs.texcoords.push_back(mesh.t[2 * i + 1]);
            if (i % 3 == 0) hs.materialIds.push_back(inst.materialId);
            else if (i % 3 == 2 && material.type == 4) {
                // scene.cpp:176-180 (sic): the instance-local index i addresses the scene-wide array
                float area = triangleArea(hs.vertices[i - 2], hs.vertices[i - 1], hs.vertices[i]);
                hs.lightPowerFromFile.push_back(powerUnitArea * area);
            }
        }
    }
    hs.T = (int)hs.materialIds.size();
    if (hs.T == 0) { err = "No mesh data loaded (scene.cpp:192-195)"; return false; }
    if ((int)hs.vertices.size() != 3 * hs.T) { err = "OBJ vertex count is not a multiple of 3"; return false; }
    return true;
}

}  // namespace rs
