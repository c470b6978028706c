#pragma once
#include "srl_shim_ext.h"
