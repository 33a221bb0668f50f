// Shim: BOOST_FOREACH as a range-for (used by engine/db/index/knn/*.hpp).
#pragma once
#define BOOST_FOREACH(decl, col) for (decl : col)
