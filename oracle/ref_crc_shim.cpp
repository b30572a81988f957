// Thin C shim over the reference's own vendored CRC++ (included from /root/reference at build
// time, never copied).  Exposes exactly the call the reference makes at
// aux_separator.h:579 and :679-681: CRC::Calculate(data, n, CRC::CRC_16_CCITTFALSE()).
// TEST INFRASTRUCTURE ONLY.
#include <cstddef>
#include <cstdint>
#include "CRC.h"
extern "C" uint16_t ref_crc16_ccitt_false(const uint8_t* data, size_t n) {
    return CRC::Calculate(data, n, CRC::CRC_16_CCITTFALSE());
}
