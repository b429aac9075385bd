// Host C++ mirror: `Universe` and its factories (reference hdr/universe.h:12-109).
#pragma once
#include <cassert>
#include <iostream>
#include <string>
#include <utility>

#include "ipic3d/app/cell.h"
#include "ipic3d/app/field.h"
#include "ipic3d/app/universe_properties.h"

namespace ipic3d {

	struct Universe {
		const UniverseProperties properties;
		Cells cells;                       // size^3
		Field field;                       // (size + 3)^3: nodes with one ghost layer
		BcField bcfield;                   // (size + 2)^3: cell centres with ghost layer
		CurrentDensity currentDensity;     // (size + 1)^3: nodes

		// an empty universe of the given shape
		explicit Universe(const UniverseProperties& properties = UniverseProperties())
			: Universe(properties, Cells(properties.size), Field(properties.size + coordinate_type(3)),
			           BcField(properties.size + coordinate_type(2)), CurrentDensity(properties.size + coordinate_type(1))) {}

		// takes ownership of grids built elsewhere; their shapes must fit the properties
		Universe(const UniverseProperties& properties, Cells&& cs, Field&& f, BcField&& bcf, CurrentDensity&& cD)
			: properties(properties), cells(std::move(cs)), field(std::move(f)), bcfield(std::move(bcf)), currentDensity(std::move(cD)) {
			assert(shapesFit());
		}

		bool shapesFit() const {
			const coordinate_type& n = properties.size;
			return cells.size() == n && field.size() == n + coordinate_type(3) && bcfield.size() == n + coordinate_type(2) &&
			       currentDensity.size() == n + coordinate_type(1);
		}
	};

	namespace detail {
		// the grids every factory adds to a set of cells: dipole field on the nodes, its cell-centred average, J = 0
		inline Universe assembleUniverse(const UniverseProperties& properties, const InitProperties& initProperties, Cells&& cells) {
			Field field = initFields(initProperties, properties);
			BcField bcField = initBcFields(properties, field);
			return Universe(properties, std::move(cells), std::move(field), std::move(bcField), initCurrentDensity(properties));
		}
	}

	// `ipic3d <deck>` as the decks describe it (maxwellian per cell; reference universe.h:56-80)
	inline Universe createUniverseFromParams(const Parameters& params, const std::string& baseName) {
		const InitProperties initProperties(params);
		UniverseProperties universeProperties(params);
		universeProperties.outputFileBaseName = baseName;
		std::cout << initProperties << universeProperties;
		return detail::assembleUniverse(universeProperties, initProperties, initCells(params, initProperties, universeProperties));
	}

	// numParticles particles drawn from a generator and sorted into cells (reference universe.h:85-105)
	template<typename Distribution>
	Universe createUniverseFromDistribution(const UniverseProperties& universeProperties, const InitProperties& initProperties,
	                                        std::uint64_t numParticles, const Distribution& distribution) {
		return detail::assembleUniverse(universeProperties, initProperties, initCells(universeProperties, numParticles, distribution));
	}

	inline std::uint64_t countParticlesInDomain(const Universe& universe) { return countParticlesInDomain(universe.cells); }

} // namespace ipic3d
