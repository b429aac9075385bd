// TEST INFRASTRUCTURE ONLY (oracle/). Stand-in for allscale/utils/printer/vectors.h:
// stream printing of std::vector as "[a,b,...]".
#pragma once
#include <ostream>
#include <vector>
namespace std {
	template<typename T, typename A>
	ostream& operator<<(ostream& out, const vector<T,A>& v) {
		out << "[";
		for (size_t i = 0; i < v.size(); ++i) { if (i) out << ","; out << v[i]; }
		return out << "]";
	}
}
