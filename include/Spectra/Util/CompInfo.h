// B200 shim of Spectra/Util/CompInfo.h:17-30.
#ifndef SPECTRA_B200_COMP_INFO_H
#define SPECTRA_B200_COMP_INFO_H

namespace Spectra {

enum class CompInfo
{
    Successful,
    NotComputed,
    NotConverging,
    NumericalIssue
};

}  // namespace Spectra
#endif
