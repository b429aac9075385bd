// Host C++ mirror: the `.inp` parser (reference hdr/parameters.h:145-381).
//
// Behaviour kept exactly: line based; text after the first '#' is dropped; lines
// without '=' are skipped; a line belongs to the FIRST key of the table below that occurs in it as a
// SUBSTRING (so the order of the table is part of the format); scalar values are the last
// whitespace-separated token, per-species lists are all tokens after "key ="; the last line is ignored
// unless it ends in a newline; a missing file prints to stderr and exits with EXIT_FAILURE; keys that are
// not in the table (c, th, Smooth, boundary conditions, ...) are ignored and `c` stays 1.0.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "ipic3d/app/vector.h"

namespace ipic3d {

	using dvector = std::vector<double>;
	using ivector = std::vector<int>;

	enum class UseCase { Dipole, Test };

	struct Parameters {
		double c = 1.0;                         // light speed (never read from the file)
		double dt;
		std::uint64_t ncycles;
		Vector3<double> L;                      // simulation box length
		Vector3<double> objectCenter;
		double planetRadius;                    // key "L_square"
		Vector3<int> ncells;
		Vector3<double> dspace;                 // L / ncells
		int ns;                                 // number of species
		ivector npcelx, npcely, npcelz;
		dvector qom, rhoInit, uth, vth, wth, u0, v0, w0;
		UseCase useCase;
		std::string wmethod, SimName, PoissonCorrection, SaveDirName;
		double delta;
		Vector3<double> B0, B1;
		int FieldOutputCycle;
		std::string FieldOutputTag, MomentsOutputTag;
		int ParticlesOutputCycle;
		std::string ParticlesOutputTag;

		static std::vector<std::string> split(const std::string& line) {
			// tokens between runs of whitespace; a leading run yields an empty first token, exactly like the
			// reference's regex split on "\\s+" (the first token is never used)
			std::vector<std::string> out;
			std::size_t i = 0;
			const auto isSpace = [](char ch) { return ch == ' ' || ch == '\t' || ch == '\r' || ch == '\n' || ch == '\f' || ch == '\v'; };
			if (!line.empty() && isSpace(line[0])) out.emplace_back();
			while (i < line.size()) {
				while (i < line.size() && isSpace(line[i])) ++i;
				if (i >= line.size()) break;
				std::size_t j = i;
				while (j < line.size() && !isSpace(line[j])) ++j;
				out.emplace_back(line.substr(i, j - i));
				i = j;
			}
			return out;
		}

		explicit Parameters(const std::string& inputfile) {
			std::ifstream in(inputfile);
			if (!in) {
				std::cerr << "File not found: " << inputfile << std::endl;
				std::exit(EXIT_FAILURE);
			}

			enum Kind { Real, UInt64, Int, Text, Case, RealList, IntList };
			struct Key { const char* name; Kind kind; void* target; };
			const Key table[] = {
				{ "dt", Real, &dt }, { "ncycles", UInt64, &ncycles },
				{ "Lx", Real, &L.x }, { "Ly", Real, &L.y }, { "Lz", Real, &L.z },
				{ "x_center", Real, &objectCenter.x }, { "y_center", Real, &objectCenter.y }, { "z_center", Real, &objectCenter.z },
				{ "L_square", Real, &planetRadius }, { "delta", Real, &delta },
				{ "nxc", Int, &ncells.x }, { "nyc", Int, &ncells.y }, { "nzc", Int, &ncells.z },
				{ "B0x", Real, &B0.x }, { "B0y", Real, &B0.y }, { "B0z", Real, &B0.z },
				{ "B1x", Real, &B1.x }, { "B1y", Real, &B1.y }, { "B1z", Real, &B1.z },
				{ "Case", Case, &useCase }, { "SaveDirName", Text, &SaveDirName }, { "PoissonCorrection", Text, &PoissonCorrection },
				{ "WriteMethod", Text, &wmethod }, { "SimulationName", Text, &SimName }, { "ns = ", Int, &ns },
				{ "npcelx", IntList, &npcelx }, { "npcely", IntList, &npcely }, { "npcelz", IntList, &npcelz },
				{ "qom", RealList, &qom }, { "rhoINIT", RealList, &rhoInit },
				{ "uth", RealList, &uth }, { "vth", RealList, &vth }, { "wth", RealList, &wth },
				{ "u0", RealList, &u0 }, { "v0", RealList, &v0 }, { "w0", RealList, &w0 },
				{ "FieldOutputCycle", Int, &FieldOutputCycle }, { "FieldOutputTag", Text, &FieldOutputTag },
				{ "MomentsOutputTag", Text, &MomentsOutputTag }, { "ParticlesOutputCycle", Int, &ParticlesOutputCycle },
				{ "ParticlesOutputTag", Text, &ParticlesOutputTag },
			};

			std::string line;
			while (!std::getline(in, line).eof()) {
				line = line.substr(0, line.find('#'));
				if (line.find('=') == std::string::npos) continue;
				for (const Key& key : table) {
					if (line.find(key.name) == std::string::npos) continue;
					const std::vector<std::string> tok = split(line);
					switch (key.kind) {
						case Real:   *static_cast<double*>(key.target) = std::stod(tok.back()); break;
						case UInt64: *static_cast<std::uint64_t*>(key.target) = std::stoull(tok.back()); break;
						case Int:    *static_cast<int*>(key.target) = std::stoi(tok.back()); break;
						case Text:   *static_cast<std::string*>(key.target) = tok.back(); break;
						case Case:   useCase = tok.back() == "Dipole" ? UseCase::Dipole : UseCase::Test; break;
						case RealList: for (std::size_t i = 2; i < tok.size(); ++i) static_cast<dvector*>(key.target)->push_back(std::stod(tok[i])); break;
						case IntList:  for (std::size_t i = 2; i < tok.size(); ++i) static_cast<ivector*>(key.target)->push_back(std::stoi(tok[i])); break;
					}
					break;                                  // first matching key wins
				}
			}

			dspace.x = L.x / ncells.x;
			dspace.y = L.y / ncells.y;
			dspace.z = L.z / ncells.z;
		}
	};

} // namespace ipic3d
