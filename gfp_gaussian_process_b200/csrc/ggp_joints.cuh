// ggp_joints.cuh — placeholder until the joints kernels land
#pragma once
