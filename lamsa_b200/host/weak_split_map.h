#pragma weak init_hash
#pragma weak hash_split_map
