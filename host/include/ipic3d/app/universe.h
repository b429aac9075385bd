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
		Cells cells;
		Field field;
		BcField bcfield;
		CurrentDensity currentDensity;

		explicit Universe(const UniverseProperties& properties = UniverseProperties())
			: properties(properties), cells(properties.size), field(properties.size + coordinate_type(3)),
			  bcfield(properties.size + coordinate_type(2)), currentDensity(properties.size + coordinate_type(1)) {}

		Universe(const UniverseProperties& properties, Cells&& cs, Field&& f, BcField&& bcf, CurrentDensity&& cD)
			: properties(properties), cells(std::move(cs)), field(std::move(f)), bcfield(std::move(bcf)), currentDensity(std::move(cD)) {
			assert(cells.size() == properties.size);
			assert(field.size() == properties.size + coordinate_type(3));
			assert(bcfield.size() == properties.size + coordinate_type(2));
			assert(currentDensity.size() == properties.size + coordinate_type(1));
		}
	};

	inline Universe createUniverseFromParams(const Parameters& params, const std::string& baseName) {
		const InitProperties initProperties(params);
		std::cout << initProperties;
		UniverseProperties universeProperties(params);
		universeProperties.outputFileBaseName = baseName;
		std::cout << universeProperties;
		Cells cells = initCells(params, initProperties, universeProperties);
		Field field = initFields(initProperties, universeProperties);
		BcField bcField = initBcFields(universeProperties, field);
		return Universe(universeProperties, std::move(cells), std::move(field), std::move(bcField), initCurrentDensity(universeProperties));
	}

	template<typename Distribution>
	Universe createUniverseFromDistribution(const UniverseProperties& universeProperties, const InitProperties& initProperties,
	                                        std::uint64_t numParticles, const Distribution& distribution) {
		Cells cells = initCells(universeProperties, numParticles, distribution);
		Field field = initFields(initProperties, universeProperties);
		BcField bcField = initBcFields(universeProperties, field);
		return Universe(universeProperties, std::move(cells), std::move(field), std::move(bcField), initCurrentDensity(universeProperties));
	}

	inline std::uint64_t countParticlesInDomain(const Universe& universe) { return countParticlesInDomain(universe.cells); }

} // namespace ipic3d
