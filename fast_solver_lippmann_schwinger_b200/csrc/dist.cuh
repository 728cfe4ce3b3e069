// NCCL plumbing shared by the sharded 3-D operator and the Krylov reductions.
#pragma once
#include <nccl.h>
#include <cstring>
