// TEST INFRASTRUCTURE (CPU oracle) - never linked into the product.
//
// Enumerations, wire structs and constant tables of the reference's generation path:
//   Block            /root/reference/src/terrain/block.hpp:5-154
//   Biome ... CaveFeaturePlacement   /root/reference/src/terrain/biome.hpp:13-260
//   tables filled by BiomeUtils::init()  /root/reference/src/terrain/biomeFuncs.hpp:725-1256
// The tables are written out as plain initialisers (values and order follow init()).
#pragma once
#include <cstdint>

namespace mmo {

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
static const uint8_t kBiomeNoiseWeights[NUM_BIOMES][6] = {
    {1, 2, 1, 1, 0, 0}, {1, 2, 1, 2, 0, 0}, {1, 2, 2, 0, 1, 0}, {1, 2, 2, 1, 2, 0}, {1, 2, 2, 2, 2, 0},
    {1, 1, 1, 0, 0, 0}, {1, 1, 2, 0, 1, 0}, {1, 1, 2, 0, 2, 0},
    {2, 0, 1, 1, 1, 1}, {2, 0, 1, 1, 1, 2}, {2, 0, 1, 1, 2, 1}, {2, 0, 1, 1, 2, 2},
    {2, 0, 1, 2, 1, 1}, {2, 0, 1, 2, 1, 2}, {2, 0, 1, 2, 2, 1}, {2, 0, 1, 2, 2, 2},
    {2, 0, 2, 1, 1, 1}, {2, 0, 2, 1, 1, 2}, {2, 0, 2, 1, 2, 1}, {2, 0, 2, 1, 2, 2},
    {2, 0, 2, 2, 1, 1}, {2, 0, 2, 2, 1, 2}, {2, 0, 2, 2, 2, 1}, {2, 0, 2, 2, 2, 2}};
// cave biome table, biomeFuncs.hpp:767-776; columns: none, shallow, warped, rocky
static const uint8_t kCaveBiomeNoiseWeights[NUM_CAVE_BIOMES][4] = {
    {1, 0, 0, 0}, {2, 1, 0, 1}, {2, 1, 0, 2}, {0, 2, 1, 0}, {0, 2, 2, 0}};

// grass block per biome, biomeFuncs.hpp:786-801 (default DIRT)
static const uint8_t kBiomeGrassBlock[NUM_BIOMES] = {
    B_DIRT, B_DIRT, B_DIRT, B_DIRT, B_DIRT, B_DIRT, B_JUNGLE_GRASS_BLOCK, B_DIRT,
    B_SAVANNA_GRASS_BLOCK, B_DIRT, B_SNOWY_GRASS_BLOCK, B_GRASS_BLOCK, B_JUNGLE_GRASS_BLOCK, B_DIRT, B_GRASS_BLOCK, B_GRASS_BLOCK,
    B_JUNGLE_GRASS_BLOCK, B_DIRT, B_MYCELIUM, B_DIRT, B_JUNGLE_GRASS_BLOCK, B_DIRT, B_GRASS_BLOCK, B_GRASS_BLOCK};

// material infos, biomeFuncs.hpp:808-847: block, thickness, noise amplitude | tan(angle of repose),
// noise scale | max slope. The eroded materials' second value is tanf(radians(angle)) evaluated on
// the HOST by the reference (glibc here); the bit patterns below are what that evaluates to with
// this image's glibc for 55,40,45,40,30,35,65,45 degrees (checked by tests/test_tables.py).
struct MaterialInfo { uint8_t block; float thickness, v1, v2; };
float material_tan_repose(int erodedIdx);   // defined in mm_tables.cpp
const MaterialInfo* material_infos();        // 20 entries

// biome -> material weights [biome][material], biomeFuncs.hpp:856-962
const float* biome_material_weights();       // 24*20, index material + 20*biome

// dirVecs2d, util/enums.hpp:32-41 (N, NE, E, SE, S, SW, W, NW)
static const int kDirVecs2d[8][2] = {{0, 1}, {1, 1}, {1, 0}, {1, -1}, {0, -1}, {-1, -1}, {-1, 0}, {-1, 1}};

// feature height bounds, biomeFuncs.hpp:1042-1074 and 1210-1223
static const int kFeatureHeightBounds[NUM_FEATURES][2] = {
    {0, 0}, {-6, 6}, {-3, 12}, {0, 20}, {0, 110}, {0, 15}, {-5, 75}, {-3, 50}, {0, 30}, {0, 15}, {0, 8}, {0, 10},
    {0, 38}, {0, 17}, {0, 5}, {0, 6}, {0, 120}, {-3, 32}, {-6, 64}, {0, 28}, {0, 15}};
static const int kCaveFeatureHeightBounds[NUM_CAVE_FEATURES][2] = {
    {0, 0}, {-3, 3}, {-3, 3}, {0, 0}, {0, 6}, {-12, 12}, {-12, 12}, {-8, 8}, {-2, 3}, {-2, 5}};

// feature generators (biomeFuncs.hpp:975-1040): per biome, in order
struct TopLayer { uint8_t material; float minThickness; };
struct FeatureGen { uint8_t feature; int gridCellSize, gridCellPadding; float chance; int numTop; TopLayer top[2]; bool canReplace; };
struct CaveFeatureGen { uint8_t feature; int gridCellSize, gridCellPadding; float chance; int minLayerHeight; bool canReplace, fromCeiling, inLava; };
struct DecoratorGen { uint8_t block; float chance; int numUnder; uint8_t under[3]; uint8_t replace; uint8_t second; bool fromCeiling; };
struct GenList { int n; const void* gens; };
const FeatureGen* biome_feature_gens(int biome, int* n);
const CaveFeatureGen* cave_biome_feature_gens(int caveBiome, int* n);
const DecoratorGen* biome_decorator_gens(int biome, int* n);
const DecoratorGen* cave_biome_decorator_gens(int caveBiome, int* n);

}  // namespace mmo
