#pragma once
#include <string>
#include "frame_parser.h"
#include "numeric.h"

namespace jxlb {

// Fills cp from the image's colour encoding; returns kParseOk or kParseUnsupported.
int MakeColorParams(const ImageMetadata& md, ColorParams* cp, std::string* err);

}  // namespace jxlb
