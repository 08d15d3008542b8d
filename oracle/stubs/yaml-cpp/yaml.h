// oracle/stubs — empty stand-in so the reference's SBA.cu (which reaches yaml-cpp only through
// CUDASolverBundling.h's member declaration) compiles without the yaml-cpp package. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <memory>
#include <string>
namespace YAML { class Node { public: template <class K> Node operator[](const K&) const { return Node(); } template <class T> T as() const { return T(); } }; }
