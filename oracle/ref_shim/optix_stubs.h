// Shim: empty stand-in for <optix_stubs.h> (/root/reference/src/util/common.h:5).
#pragma once
