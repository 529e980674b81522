// Minimal stand-in for MATLAB's C++ MEX API ("mex.hpp"), just enough to compile
// the reference's UNMODIFIED priority_queue_interface_mex.cpp outside MATLAB.
// TEST INFRASTRUCTURE (oracle/): lets tests pin the oracle's heap restatement
// against the reference's own native code + this container's libstdc++.
// Written from the usage in that file (inputs[i][j], TypedArray<double>,
// ArrayFactory::createScalar<T>, ArgumentList::size); not MATLAB code.
#pragma once
#include <cstddef>
#include <vector>

namespace matlab {
namespace data {

struct Elem {
    double v;
    template <class T> operator T() const { return static_cast<T>(v); }
};

class Array {
public:
    std::vector<double> values;
    bool is_signed_int = false;  // createScalar<int>(-1) marks "empty queue"
    Elem operator[](std::size_t i) const { return Elem{values[i]}; }
    std::size_t getNumberOfElements() const { return values.size(); }
};

template <class T> class TypedArray : public Array {
public:
    TypedArray() = default;
    TypedArray(const Array &a) : Array(a) {}
};

class ArrayFactory {
public:
    template <class T> TypedArray<T> createScalar(T x) {
        TypedArray<T> a;
        a.values.push_back(static_cast<double>(x));
        a.is_signed_int = (static_cast<T>(-1) < static_cast<T>(0));
        return a;
    }
};

}  // namespace data

namespace mex {

class ArgumentList {
public:
    explicit ArgumentList(std::vector<data::Array> &v) : v_(v) {}
    data::Array &operator[](std::size_t i) { return v_[i]; }
    std::size_t size() const { return v_.size(); }
private:
    std::vector<data::Array> &v_;
};

class Function {
public:
    virtual ~Function() = default;
};

}  // namespace mex
}  // namespace matlab
