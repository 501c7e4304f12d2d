// include/ogmaneo/Helpers.h — B200-native drop-in for the reference's <ogmaneo/Helpers.h>: the API types
// (Int2/Int3/IntBuffer/CircleBuffer, Helpers.h:24-139), geometry helpers (project/address*, :205-259), stream
// (de)serialisers (:311-341) and the kernel-executor entry points (:143-166). Same names, same meaning.
//
// What differs: the hierarchy classes never run their per-column work through runKernelN; that work is CUDA kernels
// (include/ogn_kernels_api.h). runKernelN are kept as plain serial host utilities with the reference's tile order and
// RNG consumption, because (a) user code may call them and (b) the Actor's RNG replay is defined in terms of them.
#pragma once

#include "SparseMatrix.h"

#include <functional>
#include <istream>
#include <ostream>
#include <random>
#include <vector>

namespace ogmaneo {
class ComputeSystem;

template <typename T> struct Vec2 {
    T x, y;
    Vec2() {}
    Vec2(T x, T y) : x(x), y(y) {}
};

// NB: sizeof(Vec3<int>) == 16 in the reference (a fourth `pad` member) and Int3 is written raw into save files
// (SURVEY.md F7), so the member must stay. Unlike the reference it is zero-initialised here.
template <typename T> struct Vec3 {
    T x, y, z;
    T pad;
    Vec3() : pad() {}
    Vec3(T x, T y, T z) : x(x), y(y), z(z), pad() {}
};

template <typename T> struct Vec4 {
    T x, y, z, w;
    Vec4() {}
    Vec4(T x, T y, T z, T w) : x(x), y(y), z(z), w(w) {}
};

typedef Vec2<int> Int2;
typedef Vec3<int> Int3;
typedef Vec4<int> Int4;
typedef Vec2<float> Float2;
typedef Vec3<float> Float3;
typedef Vec4<float> Float4;

typedef std::vector<int> IntBuffer;     // a CSDR: one active-cell index per column, column (x,y) at y + x*sizeY
typedef std::vector<float> FloatBuffer;

// Fixed-capacity ring; pushFront() makes a new logical element 0 by stepping `start` backwards.
template <typename T> struct CircleBuffer {
    std::vector<T> data;
    int start;

    CircleBuffer() : start(0) {}
    void resize(int size) { data.resize(size); }
    void pushFront() {
        if (--start < 0) start += (int)data.size();
    }
    T &front() { return data[start]; }
    const T &front() const { return data[start]; }
    T &back() { return data[(start + data.size() - 1) % data.size()]; }
    const T &back() const { return data[(start + data.size() - 1) % data.size()]; }
    T &operator[](int index) { return data[(start + index) % data.size()]; }
    const T &operator[](int index) const { return data[(start + index) % data.size()]; }
    int size() const { return (int)data.size(); }
};

// --- kernel executors (host utilities; tile order and RNG draws as Helpers.cpp:15-110) ---
void runKernel1(ComputeSystem &cs, const std::function<void(int, std::mt19937 &rng)> &func, int size, std::mt19937 &rng,
                int batchSize);
void runKernel2(ComputeSystem &cs, const std::function<void(const Int2 &, std::mt19937 &rng)> &func, const Int2 &size,
                std::mt19937 &rng, const Int2 &batchSize);
void runKernel3(ComputeSystem &cs, const std::function<void(const Int3 &, std::mt19937 &rng)> &func, const Int3 &size,
                std::mt19937 &rng, const Int3 &batchSize);

// --- basic per-element bodies ---
void fillInt(int pos, std::mt19937 &rng, IntBuffer *buffer, int fillValue);
void fillFloat(int pos, std::mt19937 &rng, FloatBuffer *buffer, float fillValue);
void copyInt(int pos, std::mt19937 &rng, const IntBuffer *src, IntBuffer *dst);
void copyFloat(int pos, std::mt19937 &rng, const FloatBuffer *src, FloatBuffer *dst);

// --- bounds, projection, addressing ---
inline bool inBounds0(const Int2 &pos, const Int2 &upperBound) {
    return pos.x >= 0 && pos.y >= 0 && pos.x < upperBound.x && pos.y < upperBound.y;
}
inline bool inBounds(const Int2 &pos, const Int2 &lowerBound, const Int2 &upperBound) {
    return pos.x >= lowerBound.x && pos.y >= lowerBound.y && pos.x < upperBound.x && pos.y < upperBound.y;
}
inline Int2 project(const Int2 &pos, const Float2 &toScalars) {
    return Int2((int)(pos.x * toScalars.x + 0.5f), (int)(pos.y * toScalars.y + 0.5f));
}
inline Int2 projectf(const Float2 &pos, const Float2 &toScalars) {
    return Int2((int)(pos.x * toScalars.x + 0.5f), (int)(pos.y * toScalars.y + 0.5f));
}
inline int address2(const Int2 &pos, const Int2 &dims) { return pos.y + dims.y * pos.x; }
inline int address3(const Int3 &pos, const Int3 &dims) { return pos.z + dims.z * (pos.y + dims.y * pos.x); }
inline int address4(const Int4 &pos, const Int4 &dims) { return pos.w + dims.w * (pos.z + dims.z * (pos.y + dims.y * pos.x)); }

// --- pointer-vector getters ---
std::vector<IntBuffer *> get(std::vector<IntBuffer> &v);
std::vector<FloatBuffer *> get(std::vector<FloatBuffer> &v);
std::vector<const IntBuffer *> constGet(const std::vector<IntBuffer> &v);
std::vector<const FloatBuffer *> constGet(const std::vector<FloatBuffer> &v);
std::vector<IntBuffer *> get(CircleBuffer<IntBuffer> &v);
std::vector<FloatBuffer *> get(CircleBuffer<FloatBuffer> &v);
std::vector<const IntBuffer *> constGet(const CircleBuffer<IntBuffer> &v);
std::vector<const FloatBuffer *> constGet(const CircleBuffer<FloatBuffer> &v);

inline float sigmoid(float x) {
    if (x < 0.0f) {
        float z = std::exp(x);
        return z / (1.0f + z);
    }
    return 1.0f / (1.0f + std::exp(-x));
}

// --- serialisation: `int count` followed by the raw elements (Helpers.h:311-341) ---
template <typename T> void writeBufferToStream(std::ostream &os, const std::vector<T> *buf) {
    int count = (int)buf->size();
    os.write(reinterpret_cast<const char *>(&count), sizeof(int));
    if (count > 0) os.write(reinterpret_cast<const char *>(buf->data()), (std::streamsize)count * sizeof(T));
}
template <typename T> void readBufferFromStream(std::istream &is, std::vector<T> *buf) {
    int count = 0;
    is.read(reinterpret_cast<char *>(&count), sizeof(int));
    buf->resize(count > 0 ? count : 0);
    if (count > 0) is.read(reinterpret_cast<char *>(buf->data()), (std::streamsize)count * sizeof(T));
}

// CSR pattern of a local receptive field (host utility; the device path uses closed-form geometry instead)
void initSMLocalRF(const Int3 &inSize, const Int3 &outSize, int radius, SparseMatrix &mat);
void writeSMToStream(std::ostream &os, const SparseMatrix &mat);
void readSMFromStream(std::istream &is, SparseMatrix &mat);
} // namespace ogmaneo
