// Enumerations, wire structs and constant-memory tables of the generation path.
// Values and order follow /root/reference/src/terrain/block.hpp:5-154, biome.hpp:13-260 and the
// tables BiomeUtils::init() uploads (/root/reference/src/terrain/biomeFuncs.hpp:725-1256); here they
// are compile-time __constant__ initialisers, so there is no init-order dependency.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mmg {

constexpr int NUM_BIOMES = 24, NUM_OCEAN_BIOMES = 5, NUM_OCEAN_BEACH_BIOMES = 8;
constexpr int NUM_CAVE_BIOMES = 5;
constexpr int NUM_MATERIALS = 20, NUM_STRATIFIED = 12, NUM_FORWARD = 10, NUM_ERODED = 8;
constexpr int MAX_CAVE_LAYERS = 32, MAX_FEATURES = 2048, MAX_CAVE_FEATURES = 4096;
constexpr int SEA_LEVEL = 128, LAVA_LEVEL = 8;
constexpr int ZONE_SIZE = 12, EROSION_SIDE = ZONE_SIZE * 2 * 16, EROSION_COLS = EROSION_SIDE * EROSION_SIDE;

enum Biome : uint8_t
{
    CORAL_REEF, ARCHIPELAGO, WARM_OCEAN, ICEBERGS, COOL_OCEAN, ROCKY_BEACH, TROPICAL_BEACH, BEACH,
    SAVANNA, MESA, FROZEN_WASTELAND, REDWOOD_FOREST, SHREKS_SWAMP, SPARSE_DESERT, LUSH_BIRCH_FOREST, TIANZI_MOUNTAINS,
    JUNGLE, RED_DESERT, PURPLE_MUSHROOMS, CRYSTALS, OASIS, DESERT, PLAINS, MOUNTAINS
};
enum CaveBiome : uint8_t { CB_NONE, CB_CRYSTAL_CAVES, CB_LUSH_CAVES, CB_WARPED_FOREST, CB_AMBER_FOREST };
enum Material : uint8_t
{
    M_BLACKSTONE, M_DEEPSLATE, M_SLATE, M_STONE, M_TUFF, M_CALCITE, M_GRANITE, M_TERRACOTTA, M_MARBLE, M_ANDESITE,
    M_RED_SANDSTONE, M_SANDSTONE,
    M_GRAVEL, M_CLAY, M_MUD, M_DIRT, M_RED_SAND, M_SAND, M_SMOOTH_SAND, M_SNOW
};
enum Feature : uint8_t
{
    F_NONE, F_SPHERE, F_CORAL, F_KELP, F_ICEBERG, F_ACACIA_TREE, F_REDWOOD_TREE, F_CYPRESS_TREE, F_BIRCH_TREE, F_PINE_TREE,
    F_PINE_SHRUB, F_RAFFLESIA, F_LARGE_JUNGLE_TREE, F_SMALL_JUNGLE_TREE, F_TINY_JUNGLE_TREE, F_MEDIUM_PURPLE_MUSHROOM,
    F_PURPLE_MUSHROOM, F_MEDIUM_CRYSTAL, F_CRYSTAL, F_PALM_TREE, F_CACTUS, NUM_FEATURES
};
enum CaveFeature : uint8_t
{
    CF_NONE, CF_TEST_GLOWSTONE_PILLAR, CF_TEST_SHROOMLIGHT_PILLAR, CF_CAVE_VINE, CF_GLOWSTONE_CLUSTER, CF_STORMLIGHT_SPHERE,
    CF_CEILING_STORMLIGHT_SPHERE, CF_CRYSTAL_PILLAR, CF_WARPED_FUNGUS, CF_AMBER_FUNGUS, NUM_CAVE_FEATURES
};

enum Block : uint8_t
{
    B_AIR, B_WATER, B_LAVA, B_CAVE_VINES_MAIN, B_CAVE_VINES_GLOW_MAIN, B_CAVE_VINES_END, B_CAVE_VINES_GLOW_END, B_GRASS,
    B_JUNGLE_GRASS, B_SAVANNA_GRASS, B_WARPED_MUSHROOM, B_WARPED_ROOTS, B_NETHER_SPROUTS, B_INFECTED_MUSHROOM, B_AMBER_ROOTS,
    B_DANDELION, B_POPPY, B_PITCHER_BOTTOM, B_PITCHER_TOP, B_CORNFLOWER, B_BLUE_ORCHID, B_ALLIUM, B_RED_TULIP, B_ORANGE_TULIP,
    B_WHITE_TULIP, B_PINK_TULIP, B_LILAC_BOTTOM, B_LILAC_TOP, B_PEONY_BOTTOM, B_PEONY_TOP, B_OXEYE_DAISY, B_LILY_OF_THE_VALLEY,
    B_JUNGLE_FERN, B_SMALL_MAGENTA_CRYSTAL, B_SMALL_CYAN_CRYSTAL, B_SMALL_GREEN_CRYSTAL, B_SMALL_PURPLE_MUSHROOM, B_DEAD_BUSH,
    B_HANGING_SMALL_MAGENTA_CRYSTAL, B_HANGING_SMALL_CYAN_CRYSTAL, B_HANGING_SMALL_GREEN_CRYSTAL, B_TALL_GRASS_BOTTOM,
    B_TALL_GRASS_TOP, B_TALL_JUNGLE_GRASS_BOTTOM, B_TALL_JUNGLE_GRASS_TOP, B_TORCHFLOWER, B_BRAIN_CORAL, B_BUBBLE_CORAL,
    B_FIRE_CORAL, B_HORN_CORAL, B_TUBE_CORAL, B_SEAGRASS, B_TALL_SEAGRASS_BOTTOM, B_TALL_SEAGRASS_TOP, B_KELP_MAIN, B_KELP_END,
    B_BEDROCK,
    B_STONE, B_DIRT, B_GRASS_BLOCK, B_SAND, B_GRAVEL, B_MYCELIUM, B_SNOW, B_SNOWY_GRASS_BLOCK, B_MUSHROOM_STEM,
    B_MUSHROOM_UNDERSIDE, B_PURPLE_MUSHROOM_CAP, B_MARBLE, B_ANDESITE, B_CALCITE, B_BLACKSTONE, B_TUFF, B_DEEPSLATE, B_GRANITE,
    B_SLATE, B_SANDSTONE, B_CLAY, B_RED_SAND, B_RED_SANDSTONE, B_MUD, B_JUNGLE_GRASS_BLOCK, B_RAFFLESIA_PETAL,
    B_RAFFLESIA_CENTER, B_RAFFLESIA_SPIKES, B_RAFFLESIA_STEM, B_JUNGLE_WOOD, B_JUNGLE_LEAVES_PLAIN, B_JUNGLE_LEAVES_FRUITS,
    B_CACTUS, B_PALM_WOOD, B_PALM_LEAVES, B_MAGENTA_CRYSTAL, B_CYAN_CRYSTAL, B_GREEN_CRYSTAL, B_SMOOTH_SAND, B_TERRACOTTA,
    B_YELLOW_TERRACOTTA, B_ORANGE_TERRACOTTA, B_PURPLE_TERRACOTTA, B_RED_TERRACOTTA, B_WHITE_TERRACOTTA, B_QUARTZ, B_ICE,
    B_PACKED_ICE, B_BLUE_ICE, B_SAVANNA_GRASS_BLOCK, B_BIRCH_WOOD, B_BIRCH_LEAVES, B_YELLOW_BIRCH_LEAVES, B_ORANGE_BIRCH_LEAVES,
    B_ACACIA_WOOD, B_ACACIA_LEAVES, B_SMOOTH_SANDSTONE, B_PINE_WOOD, B_PINE_LEAVES_1, B_PINE_LEAVES_2, B_REDWOOD_WOOD,
    B_REDWOOD_LEAVES, B_CYPRESS_WOOD, B_CYPRESS_LEAVES, B_GLOWSTONE, B_SHROOMLIGHT, B_WARPED_DEEPSLATE, B_WARPED_BLACKSTONE,
    B_MOSS, B_AMBER_DEEPSLATE, B_AMBER_BLACKSTONE, B_WARPED_STEM, B_WARPED_WART, B_AMBER_STEM, B_AMBER_WART, B_COBBLESTONE,
    B_COBBLED_DEEPSLATE, B_BRAIN_CORAL_BLOCK, B_BUBBLE_CORAL_BLOCK, B_FIRE_CORAL_BLOCK, B_HORN_CORAL_BLOCK, B_TUBE_CORAL_BLOCK,
    B_SEA_LANTERN, NUM_BLOCKS
};
static_assert(NUM_BLOCKS == 140 && B_KELP_END == 55 && B_BEDROCK == 56, "block.hpp:153-154");
constexpr int NUM_NON_SOLID_BLOCKS = B_KELP_END + 1;

// wire structs (biome.hpp:108-117, 207-212, 254-260); sizes 12 / 20 / 24 bytes
struct CaveLayer { int32_t start, end; uint8_t bottomBiome, topBiome; uint8_t pad[2]; };
struct FeaturePlacement { uint8_t feature; uint8_t pad0[3]; int32_t x, y, z; uint8_t canReplaceBlocks; uint8_t pad1[3]; };
struct CaveFeaturePlacement { uint8_t feature; uint8_t pad0[3]; int32_t x, y, z; int32_t layerHeight; uint8_t canReplaceBlocks; uint8_t pad1[3]; };
static_assert(sizeof(CaveLayer) == 12 && sizeof(FeaturePlacement) == 20 && sizeof(CaveFeaturePlacement) == 24, "wire layout");

// biome noise sign table, biomeFuncs.hpp:733-762: 0 ignore, 1 positive (w *= n), 2 negative (w *= 1-n)
// columns: ocean, beach, rocky, magic, temperature, moisture
__constant__ const uint8_t c_biomeNoiseWeights[NUM_BIOMES][6] = {
    {1, 2, 1, 1, 0, 0}, {1, 2, 1, 2, 0, 0}, {1, 2, 2, 0, 1, 0}, {1, 2, 2, 1, 2, 0}, {1, 2, 2, 2, 2, 0},
    {1, 1, 1, 0, 0, 0}, {1, 1, 2, 0, 1, 0}, {1, 1, 2, 0, 2, 0},
    {2, 0, 1, 1, 1, 1}, {2, 0, 1, 1, 1, 2}, {2, 0, 1, 1, 2, 1}, {2, 0, 1, 1, 2, 2},
    {2, 0, 1, 2, 1, 1}, {2, 0, 1, 2, 1, 2}, {2, 0, 1, 2, 2, 1}, {2, 0, 1, 2, 2, 2},
    {2, 0, 2, 1, 1, 1}, {2, 0, 2, 1, 1, 2}, {2, 0, 2, 1, 2, 1}, {2, 0, 2, 1, 2, 2},
    {2, 0, 2, 2, 1, 1}, {2, 0, 2, 2, 1, 2}, {2, 0, 2, 2, 2, 1}, {2, 0, 2, 2, 2, 2}};
// cave biome table, biomeFuncs.hpp:767-776; columns: none, shallow, warped, rocky
__constant__ const uint8_t c_caveBiomeNoiseWeights[NUM_CAVE_BIOMES][4] = {
    {1, 0, 0, 0}, {2, 1, 0, 1}, {2, 1, 0, 2}, {0, 2, 1, 0}, {0, 2, 2, 0}};

// grass block per biome, biomeFuncs.hpp:786-801 (default DIRT)
__constant__ const uint8_t c_biomeGrassBlock[NUM_BIOMES] = {
    B_DIRT, B_DIRT, B_DIRT, B_DIRT, B_DIRT, B_DIRT, B_JUNGLE_GRASS_BLOCK, B_DIRT,
    B_SAVANNA_GRASS_BLOCK, B_DIRT, B_SNOWY_GRASS_BLOCK, B_GRASS_BLOCK, B_JUNGLE_GRASS_BLOCK, B_DIRT, B_GRASS_BLOCK, B_GRASS_BLOCK,
    B_JUNGLE_GRASS_BLOCK, B_DIRT, B_MYCELIUM, B_DIRT, B_JUNGLE_GRASS_BLOCK, B_DIRT, B_GRASS_BLOCK, B_GRASS_BLOCK};


// material infos (biomeFuncs.hpp:808-847). For the 8 eroded materials v1 is tan(angle of repose):
// the reference evaluates tanf(radians(angle)) with the HOST libm (55,40,45,40,30,35,65,45 degrees);
// the bit patterns below are those values (glibc 2.39, identical on the GPU box image).
struct MaterialInfo { uint8_t block; float thickness, v1, v2; };
__constant__ const MaterialInfo c_materialInfos[NUM_MATERIALS] = {
    {B_BLACKSTONE, 32.f, 32.f, 0.0030f}, {B_DEEPSLATE, 66.f, 20.f, 0.0045f}, {B_SLATE, 6.f, 24.f, 0.0062f},
    {B_STONE, 40.f, 30.f, 0.0050f}, {B_TUFF, 24.f, 42.f, 0.0060f}, {B_CALCITE, 20.f, 30.f, 0.0040f},
    {B_GRANITE, 18.f, 36.f, 0.0034f}, {B_TERRACOTTA, 32.f, 16.f, 0.0020f}, {B_MARBLE, 28.f, 56.f, 0.0050f},
    {B_ANDESITE, 24.f, 48.f, 0.0030f},
    {B_RED_SANDSTONE, 3.0f, 2.0f, 0.0035f}, {B_SANDSTONE, 3.5f, 1.5f, 0.0025f},
    {B_GRAVEL, 2.5f, 1.42814791f, 1.8f}, {B_CLAY, 2.7f, 0.839099586f, 1.8f}, {B_MUD, 2.3f, 1.0f, 1.6f},
    {B_DIRT, 4.2f, 0.839099586f, 1.2f}, {B_RED_SAND, 3.5f, 0.577350318f, 1.5f}, {B_SAND, 3.8f, 0.700207531f, 1.4f},
    {B_SMOOTH_SAND, 4.5f, 2.14450693f, 4.0f}, {B_SNOW, 2.5f, 1.0f, 1.5f}};

// biome -> material weights [biome][material] (biomeFuncs.hpp:856-962)
__constant__ const float c_biomeMaterialWeights[NUM_BIOMES][NUM_MATERIALS] = {
    {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 0.f, 1.f, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.7f, 0.8f, 0.f},
    {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 0.f, 1.f, 1.f, 0.f, 0.f, 0.3f, 0.f, 0.f, 0.f, 0.f, 0.8f, 0.f, 0.f},
    {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 0.f, 1.f, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.7f, 0.f, 0.f},
    {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 0.f, 1.f, 1.f, 0.f, 0.f, 0.5f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f},
    {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 0.f, 1.f, 1.f, 0.f, 0.f, 0.5f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f},
    {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 0.f, 1.f, 1.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f},
    {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 0.f, 1.f, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f},
    {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 0.f, 1.f, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f},
    {1.f, 1.f, 1.f, 0.6f, 0.15f, 0.f, 0.2f, 3.2f, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f},
    {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 0.f, 1.f, 1.f, 0.f, 0.f, 0.f, 0.8f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f},
    {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 0.f, 0.f, 1.f, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.6f, 0.f, 0.f, 0.f, 1.1f},
    {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 0.f, 1.f, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f},
    {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 0.f, 1.f, 1.f, 0.f, 0.f, 0.f, 1.7f, 2.2f, 0.6f, 0.f, 0.f, 0.f, 0.f},
    {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 0.f, 2.f, 0.5f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 1.4f, 0.f},
    {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 0.f, 1.f, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f},
    {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 0.f, 1.f, 1.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f},
    {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 0.f, 1.f, 1.f, 0.f, 0.f, 0.f, 1.f, 1.f, 0.5f, 0.f, 0.f, 0.f, 0.f},
    {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 0.f, 1.f, 1.f, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f},
    {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 0.f, 1.f, 1.f, 0.f, 0.f, 0.4f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f},
    {1.f, 1.f, 1.f, 1.f, 1.f, 0.3f, 1.f, 0.f, 1.f, 1.f, 0.f, 0.f, 0.15f, 0.2f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f},
    {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 0.f, 1.f, 1.f, 0.f, 1.f, 0.f, 0.4f, 0.f, 0.6f, 0.f, 0.4f, 0.f, 0.f},
    {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 0.f, 1.f, 1.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f},
    {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 0.f, 1.f, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f},
    {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 0.f, 1.f, 1.f, 0.f, 0.f, 1.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f}
};

// dirVecs2d, util/enums.hpp:32-41 (N, NE, E, SE, S, SW, W, NW)
__constant__ const int c_dirVecs2d[8][2] = {{0, 1}, {1, 1}, {1, 0}, {1, -1}, {0, -1}, {-1, -1}, {-1, 0}, {-1, 1}};

// feature height bounds, biomeFuncs.hpp:1042-1074 and 1210-1223
__constant__ const int c_featureHeightBounds[NUM_FEATURES][2] = {
    {0, 0}, {-6, 6}, {-3, 12}, {0, 20}, {0, 110}, {0, 15}, {-5, 75}, {-3, 50}, {0, 30}, {0, 15}, {0, 8}, {0, 10},
    {0, 38}, {0, 17}, {0, 5}, {0, 6}, {0, 120}, {-3, 32}, {-6, 64}, {0, 28}, {0, 15}};
__constant__ const int c_caveFeatureHeightBounds[NUM_CAVE_FEATURES][2] = {
    {0, 0}, {-3, 3}, {-3, 3}, {0, 0}, {0, 6}, {-12, 12}, {-12, 12}, {-8, 8}, {-2, 3}, {-2, 5}};

// Horizontal reach of every feature type: place_feature / place_cave_feature return false for every
// voxel with |wx - p.x| > R or |wz - p.z| > R. Each R follows from the rasteriser's OWN first
// horizontal early-out (mm_placefeature.cuh, same tests as featurePlacement.hpp:147-1380), so culling
// by it cannot change a block:
//   SPHERE dot(pos,pos) > 25; CORAL hypot > 8; KELP / vines / test pillars fx = fz = 0;
//   ICEBERG needs end >= start <=> 1 - hd/radius >= (6 f - 2)/54 with radius < 32 and |f| = |fbm2<3>| < 2.48
//   (3 corners x 130 x max (0.5-d^2)^4 d x 0.79) => hd < 42.1; ACACIA max(|fx|,|fz|) > 15;
//   REDWOOD hypot(pos * s) > 12 with s >= 0.6; CYPRESS hypot > 12; BIRCH max > 8; PINE / PINE_SHRUB max > 6;
//   RAFFLESIA |pos| > 15; LARGE_JUNGLE hypot > 15; SMALL_JUNGLE hypot > 8; TINY_JUNGLE trunk or 1-block leaves;
//   MEDIUM_PURPLE_MUSHROOM |fx|+|fz| > 8; PURPLE_MUSHROOM hypot(pos * s) > 35 with s >= 0.5;
//   CRYSTALs max > 25; PALM |fx|+|fz| > 24; CACTUS max > 5;
//   GLOWSTONE_CLUSTER |top * s| > 6 with s >= 1; STORMLIGHT spheres dist > radius, radius < 7.5;
//   CRYSTAL_PILLAR dist > radius = 4 (2 (hr - 0.5)^2 + 0.5) <= 4; WARPED_FUNGUS stem, |fx|+|fz| = 1 lights, cap hypot > 3.7;
//   AMBER_FUNGUS |fx|+|fz| in {0, 1, 2} (stem and cap ring).
// NONE terminates a list scan (chunk.cu:1448-1451), so it is never culled.
constexpr int kReachAll = 1 << 20;
__constant__ const int c_featureReach[NUM_FEATURES] = {
    kReachAll, 5, 8, 0, 43, 15, 20, 12, 8, 6, 6, 15,
    15, 8, 1, 8, 70, 25, 25, 24, 5};
__constant__ const int c_caveFeatureReach[NUM_CAVE_FEATURES] = {kReachAll, 0, 0, 0, 6, 7, 7, 4, 3, 2};
// Vertical extent of the cave features, from the same rasterisers. The reference tests every cave feature
// against [y + lo, y + layerHeight + hi] (c_caveFeatureHeightBounds) because some hang from the ceiling;
// each type only ever fills a band anchored at the floor (y) or at the ceiling (y + layerHeight):
//   {a, aAtCeiling, b, bAtCeiling}: voxels outside [y + a (+ layerHeight), y + b (+ layerHeight)] are never filled.
//   test pillars fy in [0, lh]; CAVE_VINE ty in [-height, 0], height = (int)fma(u01, 12, 3) <= 14;
//   GLOWSTONE_CLUSTER |ty * 1.35 * s| <= 6, s >= 1 => |ty| <= 4; STORMLIGHT spheres |fy| resp. |ty| <= radius < 7.5;
//   CRYSTAL_PILLAR fy >= -8 and ty <= 8; WARPED_FUNGUS fy in [-2, height + 3], height <= 5;
//   AMBER_FUNGUS fy in [-2, height + 3], height <= 8.
// A candidate's y range is the intersection of the reference's bound and this band.
__constant__ const int c_caveFeatureBand[NUM_CAVE_FEATURES][4] = {
    {-512, 0, 512, 1}, {0, 0, 0, 1}, {0, 0, 0, 1}, {-14, 1, 0, 1}, {-4, 1, 4, 1}, {-7, 0, 7, 0}, {-7, 1, 7, 1}, {-8, 0, 8, 1}, {-2, 0, 8, 0}, {-2, 0, 11, 0}};

// feature / cave-feature / decorator generators (biomeFuncs.hpp:975-1040, 1081-1178, 1189-1252), flattened:
// c_*Range[biome] = {first, count} into the generator array.
struct FeatureGen { uint8_t feature; int cell, pad; float chance; int numTop; uint8_t topMat[2]; float topMin[2]; uint8_t canReplace; };
__constant__ const FeatureGen c_featureGens[] = {
    {2, 5, 0, 0.65f, 2, {18, 17}, {0.3f, 0.3f}, 1},
    {3, 8, 0, 0.5f, 2, {18, 17}, {0.3f, 0.3f}, 1},
    {4, 112, 6, 0.7f, 0, {0, 0}, {0.f, 0.f}, 1},
    {19, 48, 3, 0.35f, 1, {18, 0}, {0.3f, 0.f}, 1},
    {5, 36, 4, 0.3f, 1, {15, 0}, {0.5f, 0.f}, 1},
    {6, 16, 2, 0.7f, 1, {15, 0}, {0.5f, 0.f}, 1},
    {7, 18, 3, 0.6f, 2, {15, 14}, {0.5f, 0.5f}, 1},
    {8, 16, 2, 0.15f, 1, {15, 0}, {0.4f, 0.f}, 1},
    {8, 9, 2, 0.7f, 1, {15, 0}, {0.5f, 0.f}, 1},
    {9, 7, 1, 0.8f, 0, {0, 0}, {0.f, 0.f}, 0},
    {10, 6, 1, 0.8f, 0, {0, 0}, {0.f, 0.f}, 0},
    {11, 54, 6, 0.5f, 1, {15, 0}, {0.5f, 0.f}, 1},
    {12, 28, 3, 0.7f, 1, {15, 0}, {0.5f, 0.f}, 1},
    {13, 10, 2, 0.82f, 1, {15, 0}, {0.5f, 0.f}, 1},
    {14, 6, 1, 0.28f, 1, {15, 0}, {0.5f, 0.f}, 1},
    {19, 40, 3, 0.2f, 1, {16, 0}, {0.3f, 0.f}, 1},
    {20, 16, 2, 0.2f, 1, {16, 0}, {0.5f, 0.f}, 1},
    {15, 10, 2, 0.5f, 1, {15, 0}, {0.3f, 0.f}, 1},
    {16, 11, 3, 0.45f, 1, {15, 0}, {0.5f, 0.f}, 1},
    {17, 28, 6, 0.9f, 0, {0, 0}, {0.f, 0.f}, 1},
    {18, 52, 10, 0.8f, 0, {0, 0}, {0.f, 0.f}, 1},
    {19, 24, 3, 0.35f, 1, {17, 0}, {0.3f, 0.f}, 1},
    {20, 16, 2, 0.4f, 1, {17, 0}, {0.5f, 0.f}, 1},
    {19, 64, 3, 0.3f, 1, {17, 0}, {0.3f, 0.f}, 1},
    {20, 16, 2, 0.7f, 1, {17, 0}, {0.5f, 0.f}, 1},
};
__constant__ const int c_featureGenRange[NUM_BIOMES][2] = {{0, 2}, {2, 0}, {2, 0}, {2, 1}, {3, 0}, {3, 0}, {3, 1}, {4, 0}, {4, 1}, {5, 0}, {5, 0}, {5, 1}, {6, 2}, {8, 0}, {8, 1}, {9, 2}, {11, 4}, {15, 2}, {17, 2}, {19, 2}, {21, 2}, {23, 2}, {25, 0}, {25, 0}};
struct CaveFeatureGen { uint8_t feature; int cell, pad; float chance; int minLayerHeight; uint8_t canReplace, fromCeiling, inLava; };
__constant__ const CaveFeatureGen c_caveFeatureGens[] = {
    {5, 32, 4, 0.8f, 4, 1, 0, 0},
    {6, 32, 4, 0.8f, 4, 1, 1, 0},
    {7, 28, 5, 0.6f, 10, 0, 1, 0},
    {4, 24, 3, 0.6f, 16, 0, 1, 0},
    {3, 4, 0, 0.4f, 4, 0, 1, 0},
    {4, 16, 3, 0.8f, 16, 0, 1, 0},
    {8, 7, 1, 0.75f, 6, 0, 0, 0},
    {4, 18, 3, 0.75f, 16, 0, 1, 0},
    {9, 5, 1, 0.6f, 9, 0, 0, 0},
};
__constant__ const int c_caveFeatureGenRange[NUM_CAVE_BIOMES][2] = {{0, 0}, {0, 3}, {3, 2}, {5, 2}, {7, 2}};
struct DecoratorGen { uint8_t block; float chance; int numUnder; uint8_t under[3]; uint8_t replace, second, fromCeiling; };
__constant__ const DecoratorGen c_decoratorGens[] = {
    {51, 0.2f, 2, {60, 95, 0}, 1, 0, 0},
    {52, 0.04f, 2, {60, 95, 0}, 1, 53, 0},
    {46, 0.03f, 2, {60, 95, 0}, 1, 1, 0},
    {47, 0.03f, 2, {60, 95, 0}, 1, 1, 0},
    {48, 0.03f, 2, {60, 95, 0}, 1, 1, 0},
    {49, 0.03f, 2, {60, 95, 0}, 1, 1, 0},
    {50, 0.03f, 2, {60, 95, 0}, 1, 1, 0},
    {7, 0.2f, 1, {59, 0, 0}, 0, 0, 0},
    {31, 0.025f, 1, {59, 0, 0}, 0, 0, 0},
    {8, 0.1f, 1, {81, 0, 0}, 0, 0, 0},
    {9, 0.1f, 1, {106, 0, 0}, 0, 0, 0},
    {7, 0.2f, 1, {59, 0, 0}, 0, 0, 0},
    {41, 0.08f, 1, {59, 0, 0}, 0, 42, 0},
    {30, 0.04f, 1, {59, 0, 0}, 0, 0, 0},
    {31, 0.04f, 1, {59, 0, 0}, 0, 0, 0},
    {28, 0.02f, 1, {59, 0, 0}, 0, 29, 0},
    {8, 0.3f, 1, {81, 0, 0}, 0, 0, 0},
    {32, 0.05f, 1, {81, 0, 0}, 0, 0, 0},
    {19, 0.03f, 1, {81, 0, 0}, 0, 0, 0},
    {20, 0.03f, 1, {81, 0, 0}, 0, 0, 0},
    {21, 0.03f, 1, {81, 0, 0}, 0, 0, 0},
    {7, 0.3f, 1, {59, 0, 0}, 0, 0, 0},
    {28, 0.02f, 1, {59, 0, 0}, 0, 29, 0},
    {26, 0.02f, 1, {59, 0, 0}, 0, 27, 0},
    {15, 0.04f, 1, {59, 0, 0}, 0, 0, 0},
    {8, 0.4f, 1, {81, 0, 0}, 0, 0, 0},
    {43, 0.2f, 1, {81, 0, 0}, 0, 44, 0},
    {17, 0.03f, 1, {81, 0, 0}, 0, 18, 0},
    {32, 0.12f, 1, {81, 0, 0}, 0, 0, 0},
    {20, 0.04f, 1, {81, 0, 0}, 0, 0, 0},
    {37, 0.02f, 1, {78, 0, 0}, 0, 0, 0},
    {36, 0.1f, 1, {62, 0, 0}, 0, 0, 0},
    {33, 0.005f, 3, {57, 72, 70}, 0, 0, 0},
    {34, 0.005f, 3, {57, 72, 70}, 0, 0, 0},
    {35, 0.005f, 3, {57, 72, 70}, 0, 0, 0},
    {36, 0.02f, 1, {62, 0, 0}, 0, 0, 0},
    {33, 0.025f, 3, {57, 72, 70}, 0, 0, 0},
    {34, 0.025f, 3, {57, 72, 70}, 0, 0, 0},
    {35, 0.025f, 3, {57, 72, 70}, 0, 0, 0},
    {8, 0.2f, 1, {81, 0, 0}, 0, 0, 0},
    {19, 0.02f, 1, {81, 0, 0}, 0, 0, 0},
    {37, 0.03f, 1, {78, 0, 0}, 0, 0, 0},
    {7, 0.2f, 1, {59, 0, 0}, 0, 0, 0},
    {22, 0.01f, 1, {59, 0, 0}, 0, 0, 0},
    {23, 0.01f, 1, {59, 0, 0}, 0, 0, 0},
    {24, 0.01f, 1, {59, 0, 0}, 0, 0, 0},
    {25, 0.01f, 1, {59, 0, 0}, 0, 0, 0},
    {15, 0.03f, 1, {59, 0, 0}, 0, 0, 0},
    {16, 0.03f, 1, {59, 0, 0}, 0, 0, 0},
    {7, 0.05f, 1, {59, 0, 0}, 0, 0, 0},
    {31, 0.015f, 1, {59, 0, 0}, 0, 0, 0},
};
__constant__ const int c_decoratorGenRange[NUM_BIOMES][2] = {{0, 7}, {7, 2}, {9, 0}, {9, 0}, {9, 0}, {9, 0}, {9, 1}, {10, 0}, {10, 1}, {11, 0}, {11, 0}, {11, 5}, {16, 5}, {21, 0}, {21, 4}, {25, 0}, {25, 5}, {30, 1}, {31, 4}, {35, 4}, {39, 2}, {41, 1}, {42, 7}, {49, 2}};
__constant__ const DecoratorGen c_caveDecoratorGens[] = {
    {33, 0.015f, 0, {0, 0, 0}, 0, 0, 0},
    {34, 0.015f, 0, {0, 0, 0}, 0, 0, 0},
    {35, 0.015f, 0, {0, 0, 0}, 0, 0, 0},
    {38, 0.015f, 0, {0, 0, 0}, 0, 0, 1},
    {39, 0.015f, 0, {0, 0, 0}, 0, 0, 1},
    {40, 0.015f, 0, {0, 0, 0}, 0, 0, 1},
    {7, 0.1f, 1, {125, 0, 0}, 0, 0, 0},
    {41, 0.03f, 1, {125, 0, 0}, 0, 42, 0},
    {45, 0.02f, 1, {125, 0, 0}, 0, 0, 0},
    {10, 0.02f, 2, {123, 124, 0}, 0, 0, 0},
    {11, 0.06f, 2, {123, 124, 0}, 0, 0, 0},
    {12, 0.04f, 2, {123, 124, 0}, 0, 0, 0},
    {13, 0.02f, 2, {126, 127, 0}, 0, 0, 0},
    {14, 0.06f, 2, {126, 127, 0}, 0, 0, 0},
};
__constant__ const int c_caveDecoratorGenRange[NUM_CAVE_BIOMES][2] = {{0, 0}, {0, 6}, {6, 3}, {9, 3}, {12, 2}};

}  // namespace mmg
