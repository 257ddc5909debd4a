/* boost_shim.h — boost::shared_ptr / make_shared as aliases of the std ones (test infrastructure, see
 * eigen_shim.h).  The reference uses them only to hold per-thread accumulators and point clouds. */
#ifndef TSDF_ORACLE_BOOST_SHIM_H_
#define TSDF_ORACLE_BOOST_SHIM_H_
#include <memory>
#include <array>
namespace boost {
template <typename T> using shared_ptr = std::shared_ptr<T>;
template <typename T, typename... Args>
inline std::shared_ptr<T> make_shared(Args&&... args) { return std::make_shared<T>(std::forward<Args>(args)...); }
template <typename T, std::size_t N> using array = std::array<T, N>;
}
#endif
