// stand-in for realm/cuda/cuda_module.h — everything lives in ../runtime_impl.h
#pragma once
#include "../runtime_impl.h"
