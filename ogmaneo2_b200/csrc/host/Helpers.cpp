// csrc/host/Helpers.cpp — host utilities behind include/ogmaneo/Helpers.h and SparseMatrix.h.
// None of this is on the step path (that is CUDA); it serves API compatibility and the save-file layout.
#include <ogmaneo/ComputeSystem.h>

#include <algorithm>

namespace ogmaneo {

// One seed per tile drawn from `rng` in tile order, then the tile's elements in order — the serial schedule of
// Helpers.cpp:15-110 (which is what the reference executes with one OpenMP thread).
void runKernel1(ComputeSystem &, const std::function<void(int, std::mt19937 &)> &func, int size, std::mt19937 &rng, int batchSize) {
    std::uniform_int_distribution<int> seedDist(0, 999999);
    for (int first = 0; first < size; first += batchSize) {
        std::mt19937 tileRng(seedDist(rng));
        const int last = std::min(size, first + batchSize);
        for (int i = first; i < last; i++) func(i, tileRng);
    }
}

void runKernel2(ComputeSystem &, const std::function<void(const Int2 &, std::mt19937 &)> &func, const Int2 &size, std::mt19937 &rng,
                const Int2 &batchSize) {
    std::uniform_int_distribution<int> seedDist(0, 999999);
    const int nbx = (size.x + batchSize.x - 1) / batchSize.x, nby = (size.y + batchSize.y - 1) / batchSize.y;
    for (int tile = 0; tile < nbx * nby; tile++) { // tile index: x fastest
        const int x0 = (tile % nbx) * batchSize.x, y0 = (tile / nbx) * batchSize.y;
        std::mt19937 tileRng(seedDist(rng));
        for (int x = x0; x < std::min(size.x, x0 + batchSize.x); x++)
            for (int y = y0; y < std::min(size.y, y0 + batchSize.y); y++) func(Int2(x, y), tileRng);
    }
}

void runKernel3(ComputeSystem &, const std::function<void(const Int3 &, std::mt19937 &)> &func, const Int3 &size, std::mt19937 &rng,
                const Int3 &batchSize) {
    std::uniform_int_distribution<int> seedDist(0, 999999);
    const int nbx = (size.x + batchSize.x - 1) / batchSize.x, nby = (size.y + batchSize.y - 1) / batchSize.y,
              nbz = (size.z + batchSize.z - 1) / batchSize.z;
    for (int tile = 0; tile < nbx * nby * nbz; tile++) {
        const int x0 = (tile % nbx) * batchSize.x, y0 = ((tile / nbx) % nby) * batchSize.y, z0 = (tile / (nbx * nby)) * batchSize.z;
        std::mt19937 tileRng(seedDist(rng));
        for (int x = x0; x < std::min(size.x, x0 + batchSize.x); x++)
            for (int y = y0; y < std::min(size.y, y0 + batchSize.y); y++)
                for (int z = z0; z < std::min(size.z, z0 + batchSize.z); z++) func(Int3(x, y, z), tileRng);
    }
}

void fillInt(int pos, std::mt19937 &, IntBuffer *buffer, int fillValue) { (*buffer)[pos] = fillValue; }
void fillFloat(int pos, std::mt19937 &, FloatBuffer *buffer, float fillValue) { (*buffer)[pos] = fillValue; }
void copyInt(int pos, std::mt19937 &, const IntBuffer *src, IntBuffer *dst) { (*dst)[pos] = (*src)[pos]; }
void copyFloat(int pos, std::mt19937 &, const FloatBuffer *src, FloatBuffer *dst) { (*dst)[pos] = (*src)[pos]; }

namespace {
template <typename C, typename P> std::vector<P *> pointersOf(C &c) {
    std::vector<P *> out(c.size());
    for (int i = 0; i < (int)c.size(); i++) out[i] = &c[i];
    return out;
}
} // namespace
std::vector<IntBuffer *> get(std::vector<IntBuffer> &v) { return pointersOf<std::vector<IntBuffer>, IntBuffer>(v); }
std::vector<FloatBuffer *> get(std::vector<FloatBuffer> &v) { return pointersOf<std::vector<FloatBuffer>, FloatBuffer>(v); }
std::vector<const IntBuffer *> constGet(const std::vector<IntBuffer> &v) { return pointersOf<const std::vector<IntBuffer>, const IntBuffer>(v); }
std::vector<const FloatBuffer *> constGet(const std::vector<FloatBuffer> &v) { return pointersOf<const std::vector<FloatBuffer>, const FloatBuffer>(v); }
std::vector<IntBuffer *> get(CircleBuffer<IntBuffer> &v) { return pointersOf<CircleBuffer<IntBuffer>, IntBuffer>(v); }
std::vector<FloatBuffer *> get(CircleBuffer<FloatBuffer> &v) { return pointersOf<CircleBuffer<FloatBuffer>, FloatBuffer>(v); }
std::vector<const IntBuffer *> constGet(const CircleBuffer<IntBuffer> &v) { return pointersOf<const CircleBuffer<IntBuffer>, const IntBuffer>(v); }
std::vector<const FloatBuffer *> constGet(const CircleBuffer<FloatBuffer> &v) { return pointersOf<const CircleBuffer<FloatBuffer>, const FloatBuffer>(v); }

// The CSR pattern of a local receptive field (what Helpers.cpp:236-313 produces), built from per-axis field bounds with
// exact sizes up front instead of push_back. Values are zero.
void initSMLocalRF(const Int3 &inSize, const Int3 &outSize, int radius, SparseMatrix &mat) {
    const Float2 outToIn((float)inSize.x / (float)outSize.x, (float)inSize.y / (float)outSize.y);
    std::vector<int> x0(outSize.x), x1(outSize.x), y0(outSize.y), y1(outSize.y);
    for (int ox = 0; ox < outSize.x; ox++) {
        const int c = project(Int2(ox, 0), outToIn).x;
        x0[ox] = std::max(0, c - radius); x1[ox] = std::min(inSize.x - 1, c + radius);
    }
    for (int oy = 0; oy < outSize.y; oy++) {
        const int c = project(Int2(0, oy), outToIn).y;
        y0[oy] = std::max(0, c - radius); y1[oy] = std::min(inSize.y - 1, c + radius);
    }
    const int rows = outSize.x * outSize.y * outSize.z;
    mat.rows = rows;
    mat.columns = inSize.x * inSize.y * inSize.z;
    mat.rowRanges.assign((size_t)rows + 1, 0);
    size_t total = 0;
    for (int ox = 0; ox < outSize.x; ox++)
        for (int oy = 0; oy < outSize.y; oy++) {
            const int perRow = std::max(0, x1[ox] - x0[ox] + 1) * std::max(0, y1[oy] - y0[oy] + 1) * inSize.z;
            for (int oz = 0; oz < outSize.z; oz++) {
                mat.rowRanges[address3(Int3(ox, oy, oz), outSize)] = (int)total;
                total += perRow;
            }
        }
    mat.rowRanges[rows] = (int)total;
    mat.nonZeroValues.assign(total, 0.0f);
    mat.columnIndices.resize(total);
    for (int ox = 0; ox < outSize.x; ox++)
        for (int oy = 0; oy < outSize.y; oy++)
            for (int oz = 0; oz < outSize.z; oz++) {
                size_t j = (size_t)mat.rowRanges[address3(Int3(ox, oy, oz), outSize)];
                for (int ix = x0[ox]; ix <= x1[ox]; ix++)
                    for (int iy = y0[oy]; iy <= y1[oy]; iy++)
                        for (int iz = 0; iz < inSize.z; iz++) mat.columnIndices[j++] = address3(Int3(ix, iy, iz), inSize);
            }
    mat.nonZeroValueIndices.clear(); mat.columnRanges.clear(); mat.rowIndices.clear();
}

// Helpers.cpp:315-343
void writeSMToStream(std::ostream &os, const SparseMatrix &mat) {
    os.write(reinterpret_cast<const char *>(&mat.rows), sizeof(int));
    os.write(reinterpret_cast<const char *>(&mat.columns), sizeof(int));
    writeBufferToStream(os, &mat.nonZeroValues);
    writeBufferToStream(os, &mat.nonZeroValueIndices);
    writeBufferToStream(os, &mat.rowRanges);
    writeBufferToStream(os, &mat.columnIndices);
    writeBufferToStream(os, &mat.columnRanges);
    writeBufferToStream(os, &mat.rowIndices);
}

void readSMFromStream(std::istream &is, SparseMatrix &mat) {
    is.read(reinterpret_cast<char *>(&mat.rows), sizeof(int));
    is.read(reinterpret_cast<char *>(&mat.columns), sizeof(int));
    readBufferFromStream(is, &mat.nonZeroValues);
    readBufferFromStream(is, &mat.nonZeroValueIndices);
    readBufferFromStream(is, &mat.rowRanges);
    readBufferFromStream(is, &mat.columnIndices);
    readBufferFromStream(is, &mat.columnRanges);
    readBufferFromStream(is, &mat.rowIndices);
}

// --- SparseMatrix host primitives (the seven the step path's algorithm is defined with) ---------------------------------
// Transpose index: counting sort of the CSR entries by column (SparseMatrix.cpp:59-106).
void SparseMatrix::initT() {
    const size_t nnz = nonZeroValues.size();
    columnRanges.assign((size_t)columns + 1, 0);
    for (size_t j = 0; j < nnz; j++) columnRanges[(size_t)columnIndices[j] + 1]++;
    for (int c = 0; c < columns; c++) columnRanges[c + 1] += columnRanges[c];
    rowIndices.resize(nnz);
    nonZeroValueIndices.resize(nnz);
    std::vector<int> next(columnRanges.begin(), columnRanges.end() - 1);
    for (int r = 0; r < rows; r++)
        for (int j = rowRanges[r]; j < rowRanges[r + 1]; j++) {
            const int t = next[columnIndices[j]]++;
            rowIndices[t] = r;
            nonZeroValueIndices[t] = j;
        }
}

float SparseMatrix::multiplyOHVs(const std::vector<int> &nonZeroIndices, int row, int oneHotSize) const {
    float sum = 0.0f;
    for (int jj = rowRanges[row]; jj < rowRanges[row + 1]; jj += oneHotSize)
        sum += nonZeroValues[jj + nonZeroIndices[columnIndices[jj] / oneHotSize]];
    return sum;
}

float SparseMatrix::multiplyOHVsT(const std::vector<int> &nonZeroIndices, int column, int oneHotSize) const {
    float sum = 0.0f;
    for (int jj = columnRanges[column]; jj < columnRanges[column + 1]; jj += oneHotSize)
        sum += nonZeroValues[nonZeroValueIndices[jj + nonZeroIndices[rowIndices[jj] / oneHotSize]]];
    return sum;
}

void SparseMatrix::deltaOHVs(const std::vector<int> &nonZeroIndices, float delta, int row, int oneHotSize) {
    for (int jj = rowRanges[row]; jj < rowRanges[row + 1]; jj += oneHotSize)
        nonZeroValues[jj + nonZeroIndices[columnIndices[jj] / oneHotSize]] += delta;
}

void SparseMatrix::deltaChangedOHVsT(const std::vector<int> &nonZeroIndices, const std::vector<int> &nonZeroIndicesPrev, float delta,
                                     int column, int oneHotSize) {
    for (int jj = columnRanges[column]; jj < columnRanges[column + 1]; jj += oneHotSize) {
        const int i = rowIndices[jj] / oneHotSize;
        if (nonZeroIndices[i] != nonZeroIndicesPrev[i]) nonZeroValues[nonZeroValueIndices[jj + nonZeroIndices[i]]] += delta;
    }
}

} // namespace ogmaneo
