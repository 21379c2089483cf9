// Constant rule tables shared by all kernels (uploaded by fr3d_upload_tables).
#pragma once

#include "../../include/fr3d_b200.h"

namespace fr3d {

__constant__ fr3d_static_tables_t c_tab;

}  // namespace fr3d
