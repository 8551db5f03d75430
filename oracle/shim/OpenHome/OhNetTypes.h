// ohNet shim (TEST INFRASTRUCTURE ONLY): OpenHome/OhNetTypes.h is the historic name of OpenHome/Types.h
#pragma once
#include <OpenHome/Types.h>
