// TEST INFRASTRUCTURE (CPU oracle). Tables of BiomeUtils::init(),
// /root/reference/src/terrain/biomeFuncs.hpp:725-1256, as plain data.
#include "mm_tables.h"
#include "mm_devmath.h"
#include <initializer_list>

namespace mmo {

// tanf(radians(angle)) for GRAVEL..SNOW (55,40,45,40,30,35,65,45 degrees), host libm of this image
// (biomeFuncs.hpp:843-847)
static const uint32_t kTanRepose[NUM_ERODED] = {0x3FB6CD8Du, 0x3F56CF3Bu, 0x3F800000u, 0x3F56CF3Bu,
                                                0x3F13CD3Bu, 0x3F3340CDu, 0x40093F9Au, 0x3F800000u};
float material_tan_repose(int erodedIdx) { return dm_u2f(kTanRepose[erodedIdx]); }

// Both tables are built inside the initialiser of a function-local static (thread-safe since C++11): the first
// callers are the worker threads of mmo_layers' parallel_for, which must never see a half-filled table.
struct MaterialTable { MaterialInfo infos[NUM_MATERIALS]; };
const MaterialInfo* material_infos()
{
    static const MaterialTable table = [] {
    MaterialTable t = {{
        // stratified: thickness, noise amplitude, noise scale (biomeFuncs.hpp:814-827)
        {B_BLACKSTONE, 32.f, 32.f, 0.0030f}, {B_DEEPSLATE, 66.f, 20.f, 0.0045f}, {B_SLATE, 6.f, 24.f, 0.0062f},
        {B_STONE, 40.f, 30.f, 0.0050f}, {B_TUFF, 24.f, 42.f, 0.0060f}, {B_CALCITE, 20.f, 30.f, 0.0040f},
        {B_GRANITE, 18.f, 36.f, 0.0034f}, {B_TERRACOTTA, 32.f, 16.f, 0.0020f}, {B_MARBLE, 28.f, 56.f, 0.0050f},
        {B_ANDESITE, 24.f, 48.f, 0.0030f},
        {B_RED_SANDSTONE, 3.0f, 2.0f, 0.0035f}, {B_SANDSTONE, 3.5f, 1.5f, 0.0025f},
        // eroded: thickness, tan(angle of repose), max slope (biomeFuncs.hpp:830-837)
        {B_GRAVEL, 2.5f, 0.f, 1.8f}, {B_CLAY, 2.7f, 0.f, 1.8f}, {B_MUD, 2.3f, 0.f, 1.6f}, {B_DIRT, 4.2f, 0.f, 1.2f},
        {B_RED_SAND, 3.5f, 0.f, 1.5f}, {B_SAND, 3.8f, 0.f, 1.4f}, {B_SMOOTH_SAND, 4.5f, 0.f, 4.0f}, {B_SNOW, 2.5f, 0.f, 1.5f}}};
    for (int i = 0; i < NUM_ERODED; ++i) t.infos[NUM_STRATIFIED + i].v1 = material_tan_repose(i);
    return t;
    }();
    return table.infos;
}

struct WeightTable { float w[NUM_BIOMES * NUM_MATERIALS]; };
const float* biome_material_weights()
{
    static const WeightTable table = [] {
    WeightTable t;
    float* w = t.w;
    auto set = [&](int biome, int material, float v) { w[material + NUM_MATERIALS * biome] = v; };
    for (int i = 0; i < NUM_BIOMES * NUM_MATERIALS; ++i) w[i] = 1.f;
    for (int b = 0; b < NUM_BIOMES; ++b)
        for (int m : {M_TERRACOTTA, M_RED_SANDSTONE, M_SANDSTONE, M_GRAVEL, M_CLAY, M_MUD, M_RED_SAND, M_SAND, M_SMOOTH_SAND, M_SNOW})
            set(b, m, 0.f);
    set(CORAL_REEF, M_DIRT, 0.0f); set(CORAL_REEF, M_SAND, 0.7f); set(CORAL_REEF, M_SMOOTH_SAND, 0.8f);
    set(ARCHIPELAGO, M_GRAVEL, 0.3f); set(ARCHIPELAGO, M_DIRT, 0.0f); set(ARCHIPELAGO, M_SAND, 0.8f);
    set(WARM_OCEAN, M_DIRT, 0.0f); set(WARM_OCEAN, M_SAND, 0.7f);
    set(ICEBERGS, M_GRAVEL, 0.5f); set(ICEBERGS, M_DIRT, 0.0f);
    set(COOL_OCEAN, M_GRAVEL, 0.5f); set(COOL_OCEAN, M_DIRT, 0.0f);
    set(ROCKY_BEACH, M_DIRT, 0.0f); set(ROCKY_BEACH, M_GRAVEL, 1.0f);
    set(TROPICAL_BEACH, M_DIRT, 0.0f); set(TROPICAL_BEACH, M_SMOOTH_SAND, 1.0f);
    set(BEACH, M_DIRT, 0.0f); set(BEACH, M_SAND, 1.0f);
    set(SAVANNA, M_STONE, 0.6f); set(SAVANNA, M_TUFF, 0.15f); set(SAVANNA, M_CALCITE, 0.0f); set(SAVANNA, M_GRANITE, 0.2f);
    set(SAVANNA, M_TERRACOTTA, 3.2f); set(SAVANNA, M_MARBLE, 0.0f);
    set(MESA, M_CLAY, 0.8f); set(MESA, M_DIRT, 0.0f);
    set(FROZEN_WASTELAND, M_GRANITE, 0.0f); set(FROZEN_WASTELAND, M_DIRT, 0.6f); set(FROZEN_WASTELAND, M_SNOW, 1.1f);
    set(SHREKS_SWAMP, M_CLAY, 1.7f); set(SHREKS_SWAMP, M_MUD, 2.2f); set(SHREKS_SWAMP, M_DIRT, 0.6f);
    set(SPARSE_DESERT, M_MARBLE, 2.0f); set(SPARSE_DESERT, M_ANDESITE, 0.5f); set(SPARSE_DESERT, M_DIRT, 0.0f);
    set(SPARSE_DESERT, M_SMOOTH_SAND, 1.4f);
    set(TIANZI_MOUNTAINS, M_SANDSTONE, 1.0f);
    set(JUNGLE, M_CLAY, 1.0f); set(JUNGLE, M_MUD, 1.0f); set(JUNGLE, M_DIRT, 0.5f);
    set(RED_DESERT, M_RED_SANDSTONE, 1.0f); set(RED_DESERT, M_DIRT, 0.0f); set(RED_DESERT, M_RED_SAND, 1.0f);
    set(PURPLE_MUSHROOMS, M_GRAVEL, 0.4f);
    set(CRYSTALS, M_CALCITE, 0.3f); set(CRYSTALS, M_GRAVEL, 0.15f); set(CRYSTALS, M_CLAY, 0.2f); set(CRYSTALS, M_DIRT, 0.0f);
    set(OASIS, M_SANDSTONE, 1.0f); set(OASIS, M_CLAY, 0.4f); set(OASIS, M_DIRT, 0.6f); set(OASIS, M_SAND, 0.4f);
    set(DESERT, M_SANDSTONE, 1.0f); set(DESERT, M_DIRT, 0.0f); set(DESERT, M_SAND, 1.0f);
    set(MOUNTAINS, M_GRAVEL, 1.0f);
    return t;
    }();
    return table.w;
}

// ---- feature generators (biomeFuncs.hpp:975-1040) ----
#define FG(f, cs, cp, ch, nt, t0m, t0t, t1m, t1t, rep) {f, cs, cp, ch, nt, {{t0m, t0t}, {t1m, t1t}}, rep}
static const FeatureGen fgCoral[] = {FG(F_CORAL, 5, 0, 0.65f, 2, M_SMOOTH_SAND, 0.3f, M_SAND, 0.3f, true),
                                     FG(F_KELP, 8, 0, 0.50f, 2, M_SMOOTH_SAND, 0.3f, M_SAND, 0.3f, true)};
static const FeatureGen fgIcebergs[] = {FG(F_ICEBERG, 112, 6, 0.70f, 0, 0, 0.f, 0, 0.f, true)};
static const FeatureGen fgTropicalBeach[] = {FG(F_PALM_TREE, 48, 3, 0.35f, 1, M_SMOOTH_SAND, 0.3f, 0, 0.f, true)};
static const FeatureGen fgSavanna[] = {FG(F_ACACIA_TREE, 36, 4, 0.3f, 1, M_DIRT, 0.5f, 0, 0.f, true)};
static const FeatureGen fgRedwood[] = {FG(F_REDWOOD_TREE, 16, 2, 0.70f, 1, M_DIRT, 0.5f, 0, 0.f, true)};
static const FeatureGen fgSwamp[] = {FG(F_CYPRESS_TREE, 18, 3, 0.6f, 2, M_DIRT, 0.5f, M_MUD, 0.5f, true),
                                     FG(F_BIRCH_TREE, 16, 2, 0.15f, 1, M_DIRT, 0.4f, 0, 0.f, true)};
static const FeatureGen fgBirch[] = {FG(F_BIRCH_TREE, 9, 2, 0.7f, 1, M_DIRT, 0.5f, 0, 0.f, true)};
static const FeatureGen fgTianzi[] = {FG(F_PINE_TREE, 7, 1, 0.80f, 0, 0, 0.f, 0, 0.f, false),
                                      FG(F_PINE_SHRUB, 6, 1, 0.80f, 0, 0, 0.f, 0, 0.f, false)};
static const FeatureGen fgJungle[] = {FG(F_RAFFLESIA, 54, 6, 0.50f, 1, M_DIRT, 0.5f, 0, 0.f, true),
                                      FG(F_LARGE_JUNGLE_TREE, 28, 3, 0.70f, 1, M_DIRT, 0.5f, 0, 0.f, true),
                                      FG(F_SMALL_JUNGLE_TREE, 10, 2, 0.82f, 1, M_DIRT, 0.5f, 0, 0.f, true),
                                      FG(F_TINY_JUNGLE_TREE, 6, 1, 0.28f, 1, M_DIRT, 0.5f, 0, 0.f, true)};
static const FeatureGen fgRedDesert[] = {FG(F_PALM_TREE, 40, 3, 0.20f, 1, M_RED_SAND, 0.3f, 0, 0.f, true),
                                         FG(F_CACTUS, 16, 2, 0.20f, 1, M_RED_SAND, 0.5f, 0, 0.f, true)};
static const FeatureGen fgPurple[] = {FG(F_MEDIUM_PURPLE_MUSHROOM, 10, 2, 0.50f, 1, M_DIRT, 0.3f, 0, 0.f, true),
                                      FG(F_PURPLE_MUSHROOM, 11, 3, 0.45f, 1, M_DIRT, 0.5f, 0, 0.f, true)};
static const FeatureGen fgCrystals[] = {FG(F_MEDIUM_CRYSTAL, 28, 6, 0.9f, 0, 0, 0.f, 0, 0.f, true),
                                        FG(F_CRYSTAL, 52, 10, 0.8f, 0, 0, 0.f, 0, 0.f, true)};
static const FeatureGen fgOasis[] = {FG(F_PALM_TREE, 24, 3, 0.35f, 1, M_SAND, 0.3f, 0, 0.f, true),
                                     FG(F_CACTUS, 16, 2, 0.40f, 1, M_SAND, 0.5f, 0, 0.f, true)};
static const FeatureGen fgDesert[] = {FG(F_PALM_TREE, 64, 3, 0.30f, 1, M_SAND, 0.3f, 0, 0.f, true),
                                      FG(F_CACTUS, 16, 2, 0.70f, 1, M_SAND, 0.5f, 0, 0.f, true)};
#undef FG

const FeatureGen* biome_feature_gens(int biome, int* n)
{
#define R(arr) do { *n = (int)(sizeof(arr) / sizeof(arr[0])); return arr; } while (0)
    switch (biome)
    {
    case CORAL_REEF: R(fgCoral);
    case ICEBERGS: R(fgIcebergs);
    case TROPICAL_BEACH: R(fgTropicalBeach);
    case SAVANNA: R(fgSavanna);
    case REDWOOD_FOREST: R(fgRedwood);
    case SHREKS_SWAMP: R(fgSwamp);
    case LUSH_BIRCH_FOREST: R(fgBirch);
    case TIANZI_MOUNTAINS: R(fgTianzi);
    case JUNGLE: R(fgJungle);
    case RED_DESERT: R(fgRedDesert);
    case PURPLE_MUSHROOMS: R(fgPurple);
    case CRYSTALS: R(fgCrystals);
    case OASIS: R(fgOasis);
    case DESERT: R(fgDesert);
    default: *n = 0; return nullptr;
    }
}

// ---- cave feature generators (biomeFuncs.hpp:1189-1208) ----
// {feature, cell, pad, chance, minLayerHeight, canReplace, fromCeiling, inLava}
static const CaveFeatureGen cfCrystal[] = {{CF_STORMLIGHT_SPHERE, 32, 4, 0.80f, 4, true, false, false},
                                           {CF_CEILING_STORMLIGHT_SPHERE, 32, 4, 0.80f, 4, true, true, false},
                                           {CF_CRYSTAL_PILLAR, 28, 5, 0.60f, 10, false, true, false}};
static const CaveFeatureGen cfLush[] = {{CF_GLOWSTONE_CLUSTER, 24, 3, 0.60f, 16, false, true, false},
                                        {CF_CAVE_VINE, 4, 0, 0.40f, 4, false, true, false}};
static const CaveFeatureGen cfWarped[] = {{CF_GLOWSTONE_CLUSTER, 16, 3, 0.80f, 16, false, true, false},
                                          {CF_WARPED_FUNGUS, 7, 1, 0.75f, 6, false, false, false}};
static const CaveFeatureGen cfAmber[] = {{CF_GLOWSTONE_CLUSTER, 18, 3, 0.75f, 16, false, true, false},
                                         {CF_AMBER_FUNGUS, 5, 1, 0.60f, 9, false, false, false}};

const CaveFeatureGen* cave_biome_feature_gens(int caveBiome, int* n)
{
    switch (caveBiome)
    {
    case CB_CRYSTAL_CAVES: R(cfCrystal);
    case CB_LUSH_CAVES: R(cfLush);
    case CB_WARPED_FOREST: R(cfWarped);
    case CB_AMBER_FOREST: R(cfAmber);
    default: *n = 0; return nullptr;
    }
}

// ---- decorators (biomeFuncs.hpp:1078-1178, 1228-1252) ----
// {block, chance, numUnder, {under...}, replace, second, fromCeiling}; replace = the single
// possibleReplaceBlocks entry (AIR by default, WATER after setWater()); numUnder == 0 = any solid
#define DG(b, ch, nu, u0, u1, u2, rep, sec, ceil) {b, ch, nu, {u0, u1, u2}, rep, sec, ceil}
static const DecoratorGen dgCoral[] = {
    DG(B_SEAGRASS, 0.200f, 2, B_SAND, B_SMOOTH_SAND, 0, B_WATER, B_AIR, false),
    DG(B_TALL_SEAGRASS_BOTTOM, 0.040f, 2, B_SAND, B_SMOOTH_SAND, 0, B_WATER, B_TALL_SEAGRASS_TOP, false),
    DG(B_BRAIN_CORAL, 0.030f, 2, B_SAND, B_SMOOTH_SAND, 0, B_WATER, B_WATER, false),
    DG(B_BUBBLE_CORAL, 0.030f, 2, B_SAND, B_SMOOTH_SAND, 0, B_WATER, B_WATER, false),
    DG(B_FIRE_CORAL, 0.030f, 2, B_SAND, B_SMOOTH_SAND, 0, B_WATER, B_WATER, false),
    DG(B_HORN_CORAL, 0.030f, 2, B_SAND, B_SMOOTH_SAND, 0, B_WATER, B_WATER, false),
    DG(B_TUBE_CORAL, 0.030f, 2, B_SAND, B_SMOOTH_SAND, 0, B_WATER, B_WATER, false)};
static const DecoratorGen dgArchipelago[] = {DG(B_GRASS, 0.200f, 1, B_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false),
                                             DG(B_LILY_OF_THE_VALLEY, 0.025f, 1, B_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false)};
static const DecoratorGen dgTropicalBeach[] = {DG(B_JUNGLE_GRASS, 0.1f, 1, B_JUNGLE_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false)};
static const DecoratorGen dgSavanna[] = {DG(B_SAVANNA_GRASS, 0.1f, 1, B_SAVANNA_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false)};
static const DecoratorGen dgRedwood[] = {
    DG(B_GRASS, 0.200f, 1, B_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false),
    DG(B_TALL_GRASS_BOTTOM, 0.080f, 1, B_GRASS_BLOCK, 0, 0, B_AIR, B_TALL_GRASS_TOP, false),
    DG(B_OXEYE_DAISY, 0.040f, 1, B_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false),
    DG(B_LILY_OF_THE_VALLEY, 0.040f, 1, B_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false),
    DG(B_PEONY_BOTTOM, 0.020f, 1, B_GRASS_BLOCK, 0, 0, B_AIR, B_PEONY_TOP, false)};
static const DecoratorGen dgSwamp[] = {
    DG(B_JUNGLE_GRASS, 0.300f, 1, B_JUNGLE_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false),
    DG(B_JUNGLE_FERN, 0.050f, 1, B_JUNGLE_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false),
    DG(B_CORNFLOWER, 0.030f, 1, B_JUNGLE_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false),
    DG(B_BLUE_ORCHID, 0.030f, 1, B_JUNGLE_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false),
    DG(B_ALLIUM, 0.030f, 1, B_JUNGLE_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false)};
static const DecoratorGen dgBirch[] = {
    DG(B_GRASS, 0.300f, 1, B_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false),
    DG(B_PEONY_BOTTOM, 0.020f, 1, B_GRASS_BLOCK, 0, 0, B_AIR, B_PEONY_TOP, false),
    DG(B_LILAC_BOTTOM, 0.020f, 1, B_GRASS_BLOCK, 0, 0, B_AIR, B_LILAC_TOP, false),
    DG(B_DANDELION, 0.040f, 1, B_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false)};
static const DecoratorGen dgJungle[] = {
    DG(B_JUNGLE_GRASS, 0.400f, 1, B_JUNGLE_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false),
    DG(B_TALL_JUNGLE_GRASS_BOTTOM, 0.200f, 1, B_JUNGLE_GRASS_BLOCK, 0, 0, B_AIR, B_TALL_JUNGLE_GRASS_TOP, false),
    DG(B_PITCHER_BOTTOM, 0.030f, 1, B_JUNGLE_GRASS_BLOCK, 0, 0, B_AIR, B_PITCHER_TOP, false),
    DG(B_JUNGLE_FERN, 0.120f, 1, B_JUNGLE_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false),
    DG(B_BLUE_ORCHID, 0.040f, 1, B_JUNGLE_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false)};
static const DecoratorGen dgRedDesert[] = {DG(B_DEAD_BUSH, 0.020f, 1, B_RED_SAND, 0, 0, B_AIR, B_AIR, false)};
static const DecoratorGen dgPurple[] = {
    DG(B_SMALL_PURPLE_MUSHROOM, 0.100f, 1, B_MYCELIUM, 0, 0, B_AIR, B_AIR, false),
    DG(B_SMALL_MAGENTA_CRYSTAL, 0.005f, 3, B_STONE, B_TUFF, B_CALCITE, B_AIR, B_AIR, false),
    DG(B_SMALL_CYAN_CRYSTAL, 0.005f, 3, B_STONE, B_TUFF, B_CALCITE, B_AIR, B_AIR, false),
    DG(B_SMALL_GREEN_CRYSTAL, 0.005f, 3, B_STONE, B_TUFF, B_CALCITE, B_AIR, B_AIR, false)};
static const DecoratorGen dgCrystals[] = {
    DG(B_SMALL_PURPLE_MUSHROOM, 0.020f, 1, B_MYCELIUM, 0, 0, B_AIR, B_AIR, false),
    DG(B_SMALL_MAGENTA_CRYSTAL, 0.025f, 3, B_STONE, B_TUFF, B_CALCITE, B_AIR, B_AIR, false),
    DG(B_SMALL_CYAN_CRYSTAL, 0.025f, 3, B_STONE, B_TUFF, B_CALCITE, B_AIR, B_AIR, false),
    DG(B_SMALL_GREEN_CRYSTAL, 0.025f, 3, B_STONE, B_TUFF, B_CALCITE, B_AIR, B_AIR, false)};
static const DecoratorGen dgOasis[] = {DG(B_JUNGLE_GRASS, 0.200f, 1, B_JUNGLE_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false),
                                       DG(B_CORNFLOWER, 0.020f, 1, B_JUNGLE_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false)};
static const DecoratorGen dgDesert[] = {DG(B_DEAD_BUSH, 0.030f, 1, B_RED_SAND, 0, 0, B_AIR, B_AIR, false)};
static const DecoratorGen dgPlains[] = {
    DG(B_GRASS, 0.200f, 1, B_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false),
    DG(B_RED_TULIP, 0.010f, 1, B_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false),
    DG(B_ORANGE_TULIP, 0.010f, 1, B_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false),
    DG(B_WHITE_TULIP, 0.010f, 1, B_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false),
    DG(B_PINK_TULIP, 0.010f, 1, B_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false),
    DG(B_DANDELION, 0.030f, 1, B_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false),
    DG(B_POPPY, 0.030f, 1, B_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false)};
static const DecoratorGen dgMountains[] = {DG(B_GRASS, 0.050f, 1, B_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false),
                                           DG(B_LILY_OF_THE_VALLEY, 0.015f, 1, B_GRASS_BLOCK, 0, 0, B_AIR, B_AIR, false)};

const DecoratorGen* biome_decorator_gens(int biome, int* n)
{
    switch (biome)
    {
    case CORAL_REEF: R(dgCoral);
    case ARCHIPELAGO: R(dgArchipelago);
    case TROPICAL_BEACH: R(dgTropicalBeach);
    case SAVANNA: R(dgSavanna);
    case REDWOOD_FOREST: R(dgRedwood);
    case SHREKS_SWAMP: R(dgSwamp);
    case LUSH_BIRCH_FOREST: R(dgBirch);
    case JUNGLE: R(dgJungle);
    case RED_DESERT: R(dgRedDesert);
    case PURPLE_MUSHROOMS: R(dgPurple);
    case CRYSTALS: R(dgCrystals);
    case OASIS: R(dgOasis);
    case DESERT: R(dgDesert);
    case PLAINS: R(dgPlains);
    case MOUNTAINS: R(dgMountains);
    default: *n = 0; return nullptr;
    }
}

static const DecoratorGen dgcCrystal[] = {
    DG(B_SMALL_MAGENTA_CRYSTAL, 0.015f, 0, 0, 0, 0, B_AIR, B_AIR, false),
    DG(B_SMALL_CYAN_CRYSTAL, 0.015f, 0, 0, 0, 0, B_AIR, B_AIR, false),
    DG(B_SMALL_GREEN_CRYSTAL, 0.015f, 0, 0, 0, 0, B_AIR, B_AIR, false),
    DG(B_HANGING_SMALL_MAGENTA_CRYSTAL, 0.015f, 0, 0, 0, 0, B_AIR, B_AIR, true),
    DG(B_HANGING_SMALL_CYAN_CRYSTAL, 0.015f, 0, 0, 0, 0, B_AIR, B_AIR, true),
    DG(B_HANGING_SMALL_GREEN_CRYSTAL, 0.015f, 0, 0, 0, 0, B_AIR, B_AIR, true)};
static const DecoratorGen dgcLush[] = {DG(B_GRASS, 0.100f, 1, B_MOSS, 0, 0, B_AIR, B_AIR, false),
                                       DG(B_TALL_GRASS_BOTTOM, 0.030f, 1, B_MOSS, 0, 0, B_AIR, B_TALL_GRASS_TOP, false),
                                       DG(B_TORCHFLOWER, 0.020f, 1, B_MOSS, 0, 0, B_AIR, B_AIR, false)};
static const DecoratorGen dgcWarped[] = {
    DG(B_WARPED_MUSHROOM, 0.020f, 2, B_WARPED_DEEPSLATE, B_WARPED_BLACKSTONE, 0, B_AIR, B_AIR, false),
    DG(B_WARPED_ROOTS, 0.060f, 2, B_WARPED_DEEPSLATE, B_WARPED_BLACKSTONE, 0, B_AIR, B_AIR, false),
    DG(B_NETHER_SPROUTS, 0.040f, 2, B_WARPED_DEEPSLATE, B_WARPED_BLACKSTONE, 0, B_AIR, B_AIR, false)};
static const DecoratorGen dgcAmber[] = {
    DG(B_INFECTED_MUSHROOM, 0.020f, 2, B_AMBER_DEEPSLATE, B_AMBER_BLACKSTONE, 0, B_AIR, B_AIR, false),
    DG(B_AMBER_ROOTS, 0.060f, 2, B_AMBER_DEEPSLATE, B_AMBER_BLACKSTONE, 0, B_AIR, B_AIR, false)};
#undef DG

const DecoratorGen* cave_biome_decorator_gens(int caveBiome, int* n)
{
    switch (caveBiome)
    {
    case CB_CRYSTAL_CAVES: R(dgcCrystal);
    case CB_LUSH_CAVES: R(dgcLush);
    case CB_WARPED_FOREST: R(dgcWarped);
    case CB_AMBER_FOREST: R(dgcAmber);
    default: *n = 0; return nullptr;
    }
}
#undef R

}  // namespace mmo
