// Stub: OpenVDB is only used inside core/ResourceManager.cpp, which the oracle does not compile.
#pragma once
