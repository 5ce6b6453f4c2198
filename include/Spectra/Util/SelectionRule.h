// B200 shim of Spectra/Util/SelectionRule.h:33-58 — the enum only; sorting itself happens behind the C ABI.
#ifndef SPECTRA_B200_SELECTION_RULE_H
#define SPECTRA_B200_SELECTION_RULE_H

namespace Spectra {

enum class SortRule
{
    LargestMagn,
    LargestReal,
    LargestImag,
    LargestAlge,
    SmallestMagn,
    SmallestReal,
    SmallestImag,
    SmallestAlge,
    BothEnds
};

}  // namespace Spectra
#endif
