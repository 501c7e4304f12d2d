// include/ogmaneo/SparseMatrix.h — B200-native drop-in for the reference's <ogmaneo/SparseMatrix.h>.
//
// On the device the weights are NOT stored as CSR (see include/ogn_kernels_api.h: fp32 "F"/"R" layouts addressed by
// closed-form geometry, 64-bit offsets, no index arrays). This struct only exists so that
//   * getVisibleLayer(i).weights keeps its type and public fields (SparseCoder.h:32-36, Predictor.h:33-37,
//     Actor.h:32-35 of the reference), filled lazily from the device in the reference's value order, and
//   * save files keep the reference's byte layout (Helpers.cpp:315-343).
// Field names and meaning follow SparseMatrix.h:16-27 of the reference; only the row/column primitives the step
// path uses (count, countT, multiplyOHVs, multiplyOHVsT, deltaOHVs, deltaChangedOHVsT, initT) are
// provided as host utilities.
#pragma once

#include <vector>

namespace ogmaneo {
struct SparseMatrix {
    int rows = 0, columns = 0;

    std::vector<float> nonZeroValues;
    std::vector<int> rowRanges;
    std::vector<int> columnIndices;

    // transpose index (only SparseCoder matrices carry it)
    std::vector<int> nonZeroValueIndices;
    std::vector<int> columnRanges;
    std::vector<int> rowIndices;

    SparseMatrix() {}

    void initT();                                                        // SparseMatrix.cpp:59-106
    int count(int row) const { return rowRanges[row + 1] - rowRanges[row]; }              // :139-145
    int countT(int column) const { return columnRanges[column + 1] - columnRanges[column]; } // :215-221
    float multiplyOHVs(const std::vector<int> &nonZeroIndices, int row, int oneHotSize) const;       // :260-276
    float multiplyOHVsT(const std::vector<int> &nonZeroIndices, int column, int oneHotSize) const;   // :278-294
    void deltaOHVs(const std::vector<int> &nonZeroIndices, float delta, int row, int oneHotSize);    // :488-501
    void deltaChangedOHVsT(const std::vector<int> &nonZeroIndices, const std::vector<int> &nonZeroIndicesPrev, float delta,
                           int column, int oneHotSize);                                              // :572-590
};
} // namespace ogmaneo
