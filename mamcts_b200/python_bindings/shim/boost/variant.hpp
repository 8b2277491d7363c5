// Stand-in for boost/variant.hpp when the reference's python/bindings/common.hpp is compiled here: it only
// specialises pybind11's variant caster for boost::variant (never instantiated by the bound classes).
#pragma once
namespace boost {
template <class... Ts> class variant;
template <class V, class X> void apply_visitor(V&&, X&&);
}  // namespace boost
