// Host API, GL-free — Vec3 / Mat3 with the reference's public names and arithmetic.
//
// Mirrors src/Vec3.h of the reference (class names, member names, operator set) so scene code
// written against the reference compiles against this header. The arithmetic contracts that the
// device kernels and the KD-tree build rely on (and that tests/test_host_scene.py pins bit for
// bit against the reference) are:
//   dot(a,b)     = (a0*b0 + a1*b1) + a2*b2            (Vec3.h:36-38)
//   cross(a,b)   = (a1*b2 - a2*b1, a2*b0 - a0*b2, a0*b1 - a1*b0)   (Vec3.h:39-43)
//   normalize()  = three divisions by sqrt(squareLength())         (Vec3.h:35)
//   Mat3 * Vec3  = row . vector, accumulated left to right         (Vec3.h:150-157)
// All in fp32, never contracted (the build uses no -march, hence no FMA).
#ifndef HAI719_HOST_VEC3_H
#define HAI719_HOST_VEC3_H

#include <cmath>
#include <cstdlib>
#include <iostream>

class Vec3 {
    float mVals[3];

public:
    Vec3() : mVals{0.f, 0.f, 0.f} {}
    Vec3(float x, float y, float z) : mVals{x, y, z} {}
    Vec3(float f) : mVals{f, f, f} {}

    float &operator[](unsigned int c) { return mVals[c]; }
    float operator[](unsigned int c) const { return mVals[c]; }

    float squareLength() const { return mVals[0] * mVals[0] + mVals[1] * mVals[1] + mVals[2] * mVals[2]; }
    float length() const { return std::sqrt(squareLength()); }
    float norm() const { return length(); }
    float squareNorm() const { return squareLength(); }
    void normalize() {
        const float L = length();
        mVals[0] /= L; mVals[1] /= L; mVals[2] /= L;
    }

    static float dot(Vec3 const &a, Vec3 const &b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
    static Vec3 cross(Vec3 const &a, Vec3 const &b) {
        return Vec3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
    }
    static Vec3 compProduct(Vec3 const &a, Vec3 const &b) { return Vec3(a[0] * b[0], a[1] * b[1], a[2] * b[2]); }

    void operator+=(Vec3 const &o) { for (int i = 0; i < 3; ++i) mVals[i] += o[i]; }
    void operator-=(Vec3 const &o) { for (int i = 0; i < 3; ++i) mVals[i] -= o[i]; }
    void operator*=(float s) { for (int i = 0; i < 3; ++i) mVals[i] *= s; }
    void operator/=(float s) { for (int i = 0; i < 3; ++i) mVals[i] /= s; }

    unsigned int getMaxAbsoluteComponent() const {
        const float ax = std::fabs(mVals[0]), ay = std::fabs(mVals[1]), az = std::fabs(mVals[2]);
        if (ax > ay) return ax > az ? 0u : 2u;
        return ay > az ? 1u : 2u;
    }
    Vec3 getOrthogonal() const {
        const unsigned int c1 = getMaxAbsoluteComponent(), c2 = (c1 + 1) % 3;
        Vec3 r;
        r[c1] = mVals[c2];
        r[c2] = -mVals[c1];
        return r;
    }
};

static inline Vec3 operator+(Vec3 const &a, Vec3 const &b) { return Vec3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
static inline Vec3 operator-(Vec3 const &a, Vec3 const &b) { return Vec3(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
static inline Vec3 operator*(float a, Vec3 const &b) { return Vec3(a * b[0], a * b[1], a * b[2]); }
static inline Vec3 operator*(Vec3 const &b, float a) { return Vec3(a * b[0], a * b[1], a * b[2]); }
static inline Vec3 operator/(Vec3 const &a, float b) { return Vec3(a[0] / b, a[1] / b, a[2] / b); }
static inline std::ostream &operator<<(std::ostream &s, Vec3 const &p) { return s << p[0] << " " << p[1] << " " << p[2]; }
static inline std::istream &operator>>(std::istream &s, Vec3 &p) { return s >> p[0] >> p[1] >> p[2]; }

// Row-major 3x3:  0 1 2 / 3 4 5 / 6 7 8
class Mat3 {
    float vals[9];

public:
    Mat3() : vals{0, 0, 0, 0, 0, 0, 0, 0, 0} {}
    Mat3(float a, float b, float c, float d, float e, float f, float g, float h, float i) : vals{a, b, c, d, e, f, g, h, i} {}

    float operator()(unsigned int i, unsigned int j) const { return vals[3 * i + j]; }
    float &operator()(unsigned int i, unsigned int j) { return vals[3 * i + j]; }

    Vec3 operator*(const Vec3 &p) const {
        Vec3 r;
        for (unsigned int i = 0; i < 3; ++i) r[i] = (*this)(i, 0) * p[0] + (*this)(i, 1) * p[1] + (*this)(i, 2) * p[2];
        return r;
    }
    Mat3 operator*(const Mat3 &m) const {
        Mat3 r;
        for (unsigned int i = 0; i < 3; ++i)
            for (unsigned int j = 0; j < 3; ++j)
                r(i, j) = (*this)(i, 0) * m(0, j) + (*this)(i, 1) * m(1, j) + (*this)(i, 2) * m(2, j);
        return r;
    }
    Mat3 operator+(const Mat3 &m) const { Mat3 r; for (int k = 0; k < 9; ++k) r.vals[k] = vals[k] + m.vals[k]; return r; }
    Mat3 operator-(const Mat3 &m) const { Mat3 r; for (int k = 0; k < 9; ++k) r.vals[k] = vals[k] - m.vals[k]; return r; }
    Mat3 operator-() const { Mat3 r; for (int k = 0; k < 9; ++k) r.vals[k] = -vals[k]; return r; }
    Mat3 operator*(float s) const { Mat3 r; for (int k = 0; k < 9; ++k) r.vals[k] = vals[k] * s; return r; }
    Mat3 operator/(float s) const { Mat3 r; for (int k = 0; k < 9; ++k) r.vals[k] = vals[k] / s; return r; }
    void operator+=(const Mat3 &m) { for (int k = 0; k < 9; ++k) vals[k] += m.vals[k]; }
    void operator-=(const Mat3 &m) { for (int k = 0; k < 9; ++k) vals[k] -= m.vals[k]; }
    void operator/=(double s) { for (int k = 0; k < 9; ++k) vals[k] /= s; }

    bool isnan() const { for (int k = 0; k < 9; ++k) if (std::isnan(vals[k])) return true; return false; }
    float sqrnorm() const { float s = vals[0] * vals[0]; for (int k = 1; k < 9; ++k) s = s + vals[k] * vals[k]; return s; }
    float norm() const { return std::sqrt(sqrnorm()); }
    float determinant() const {
        return vals[0] * (vals[4] * vals[8] - vals[7] * vals[5]) - vals[1] * (vals[3] * vals[8] - vals[6] * vals[5]) +
               vals[2] * (vals[3] * vals[7] - vals[6] * vals[4]);
    }
    float trace() const { return vals[0] + vals[4] + vals[8]; }
    void transpose() { std::swap(vals[1], vals[3]); std::swap(vals[2], vals[6]); std::swap(vals[5], vals[7]); }
    Mat3 getTranspose() const { Mat3 r(*this); r.transpose(); return r; }
};

inline static Mat3 operator*(float s, const Mat3 &m) { return m * s; }
inline static std::ostream &operator<<(std::ostream &s, Mat3 const &m) {
    for (unsigned int i = 0; i < 3; ++i) s << m(i, 0) << " \t" << m(i, 1) << " \t" << m(i, 2) << std::endl;
    return s;
}
#endif
