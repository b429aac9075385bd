// Source compatibility: the AllScale API's FileIOManager (api/core/io.h) as far as the reference's code and unit tests
// use it: named text entries that are written through one stream and read back through another.  The host mirror
// writes its outputs (`<base>.out`, ConservedQuantities.out, density snapshots) as ordinary files named like the
// entry, so an entry here IS the file of that name.
#pragma once
#include <fstream>
#include <memory>
#include <string>

#include "allscale/utils/serializer.h"

namespace allscale { namespace api { namespace core {

	enum class Mode { Text, Binary };

	struct Entry { std::string name; Mode mode = Mode::Text; };

	// copyable handles onto one open file (the reference passes stream handles around by value)
	class OutputStream {
		std::shared_ptr<std::ofstream> file;
	public:
		explicit OutputStream(const Entry& e) : file(std::make_shared<std::ofstream>(e.name, e.mode == Mode::Binary ? std::ios::binary : std::ios::out)) {}
		template<typename T> OutputStream& operator<<(const T& v) { (*file) << v; return *this; }
		OutputStream& operator<<(std::ostream& (*manip)(std::ostream&)) { manip(*file); return *this; }
		template<typename T> OutputStream& write(const T& v) { file->write((const char*)&v, sizeof(T)); return *this; }
		explicit operator bool() const { return (bool)*file; }
		void close() { file->close(); }
	};
	class InputStream {
		std::shared_ptr<std::ifstream> file;
	public:
		explicit InputStream(const Entry& e) : file(std::make_shared<std::ifstream>(e.name, e.mode == Mode::Binary ? std::ios::binary : std::ios::in)) {}
		template<typename T> InputStream& operator>>(T& v) { (*file) >> v; return *this; }
		template<typename T> T read() { T v; file->read((char*)&v, sizeof(T)); return v; }
		explicit operator bool() const { return (bool)*file; }
		void close() { file->close(); }
	};

	class FileIOManager {
		FileIOManager() = default;
	public:
		static FileIOManager& getInstance() { static FileIOManager manager; return manager; }
		Entry createEntry(const std::string& name, Mode mode = Mode::Text) { return { name, mode }; }
		OutputStream openOutputStream(const Entry& e) { return OutputStream(e); }
		InputStream openInputStream(const Entry& e) { return InputStream(e); }
		void close(OutputStream& s) { s.close(); }
		void close(InputStream& s) { s.close(); }
	};

}}}
