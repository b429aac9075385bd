// TEST INFRASTRUCTURE ONLY (oracle/). Stand-in for allscale/api/user/algorithm/async.h:
// runs the task eagerly; wait()/get() are no-ops afterwards.
#pragma once
namespace allscale { namespace api { namespace user { namespace algorithm {
	namespace detail { struct completed_task { void wait() const {} void get() const {} }; }
	template<typename F> detail::completed_task async(const F& f) { f(); return {}; }
}}}}
