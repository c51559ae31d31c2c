// Umbrella header of the C++ host above the C ABI (include/ntrace_b200.h).
#pragma once
#include "ntrace/Base.hpp"
#include "ntrace/Buffer.hpp"
#include "ntrace/RayBuffer.hpp"
#include "ntrace/Scene.hpp"
#include "ntrace/CudaBVH.hpp"
#include "ntrace/CudaBVHTracer.hpp"
#include "ntrace/RayGen.hpp"
#include "ntrace/CameraControls.hpp"
#include "ntrace/Environment.hpp"
#include "ntrace/Renderer.hpp"
